"""On-disk round trips: .npz stand-ins for the reference's HDF5 dataset / K_init / log files."""
import numpy as np
import pytest
import torch as to

pytestmark = pytest.mark.gpu

if to.cuda.is_available():
    import tvo_b200
    from tvo_b200.models import BSC
    from tvo_b200.utils import io as tio
    from tvo_b200.variational import EVOVariationalStates

DEV = "cuda:0"


def test_k_init_file_npz_and_checkpoint_roundtrip(tmp_path):
    tvo_b200._set_device(to.device(DEV))
    rng = np.random.default_rng(0)
    N, S, H, D = 17, 6, 70, 9
    K0 = (rng.random((N, S, H)) < 0.2).astype(np.uint8)
    kfile = str(tmp_path / "init.npz")
    np.savez(kfile, initial_states=K0)
    st = EVOVariationalStates(N, H, S, to.float64, "randparents", "randflip", 2, 1, 1, K_init_file=kfile)
    assert np.array_equal(st.K.unpack().cpu().numpy(), K0)                      # TVOVariationalStates.py:30-38
    W = rng.standard_normal((D, H))
    m = BSC(H, D, to.tensor(W, device=DEV), to.tensor([0.7], device=DEV), to.full((H,), 0.1, device=DEV), precision=to.float64)
    Y = to.tensor(rng.standard_normal((N, D)), device=DEV)
    st.update(to.arange(N, device=DEV), Y, m)
    ck = str(tmp_path / "ckpt")  # no suffix: save and load must agree on the file
    tio.save_checkpoint(ck, m, st, extra={"epoch": np.array(3)})
    z = np.load(ck + ".npz")
    assert z["train_states"].dtype == np.uint8 and z["train_states"].shape == (N, S, H)   # reference layout on disk
    assert set(["theta/W", "theta/pies", "theta/sigma2", "train_lpj", "epoch"]) <= set(z.files)  # H5Logger's names
    K1, l1 = st.K.unpack().clone(), st.lpj.clone()
    m2 = BSC(H, D, to.zeros(D, H, device=DEV), to.tensor([1.0], device=DEV), to.full((H,), 0.5, device=DEV), precision=to.float64)
    st2 = EVOVariationalStates(N, H, S, to.float64, "randparents", "randflip", 2, 1, 1)
    W_id = m2.theta["W"].data_ptr()
    arrs = tio.load_checkpoint(ck, m2, st2)
    assert int(arrs["epoch"]) == 3 and m2.theta["W"].data_ptr() == W_id        # theta restored in place
    for k in ("W", "pies", "sigma2"):
        assert to.equal(m2.theta[k], m.theta[k])
    assert to.equal(st2.K.unpack(), K1) and to.equal(st2.lpj, l1)
    # restored model + states give the same free energy
    idx = to.arange(N, device=DEV)
    assert abs(m2.free_energy(idx, Y, st2) - m.free_energy(idx, Y, st)) < 1e-9


def test_load_shard_and_errors(tmp_path):
    tvo_b200._set_device(to.device(DEV))
    Y = np.arange(40, dtype=np.float64).reshape(10, 4)
    np.save(str(tmp_path / "y.npy"), Y)
    np.savez(str(tmp_path / "d.npz"), data=Y, other=Y[:2])
    for f in ("y.npy", "d.npz"):
        t = tio.load_shard(str(tmp_path / f), "data", dtype=to.float32)
        assert t.is_cuda and t.dtype == to.float32 and np.array_equal(t.cpu().numpy(), Y.astype(np.float32))
    with pytest.raises(KeyError):
        tio.load_shard(str(tmp_path / "d.npz"), "missing")
    with pytest.raises(KeyError):
        tio.load_initial_states(str(tmp_path / "d.npz"), 10, 1, 4)
    bad = str(tmp_path / "bad.npz")
    np.savez(bad, initial_states=np.zeros((10, 2, 4), np.uint8))
    with pytest.raises(AssertionError):
        EVOVariationalStates(10, 4, 3, to.float64, "randparents", "randflip", 2, 1, 1, K_init_file=bad)
