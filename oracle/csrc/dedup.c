/* CPU oracle (TEST INFRASTRUCTURE ONLY): C restatement of the reference's only native routine,
 * tvo/variational/_set_redundant_lpj_to_low_CPU.pyx:13-30.
 *
 * For every datapoint n and child s: if the child's H bytes equal those of a LATER child the
 * child is marked (so the last copy of a duplicate survives); otherwise, if they equal any old
 * state, it is marked.  Marking writes (float)-1e20 into new_lpj[n][s] (pyx:17 declares the
 * constant as a C float, so the fp64 variant receives -1.0000000200408773e20).
 *
 * Built by __graft_entry__.build() into oracle/_build/liboracle.so.
 */
#include <stddef.h>
#include <string.h>

#define MARK_BODY(T)                                                                           \
    const float low = -1e20f;                                                                  \
    for (size_t n = 0; n < N; ++n) {                                                           \
        const unsigned char *cn = new_states + n * Snew * H;                                   \
        const unsigned char *on = old_states + n * S * H;                                      \
        for (size_t s = 0; s < Snew; ++s) {                                                    \
            const unsigned char *c = cn + s * H;                                               \
            int dup = 0;                                                                       \
            for (size_t t = s + 1; t < Snew && !dup; ++t) dup = memcmp(c, cn + t * H, H) == 0; \
            for (size_t t = 0; t < S && !dup; ++t) dup = memcmp(c, on + t * H, H) == 0;        \
            if (dup) new_lpj[n * Snew + s] = (T)low;                                           \
        }                                                                                      \
    }

void oracle_mark_redundant_f64(const unsigned char *new_states, double *new_lpj,
                               const unsigned char *old_states, size_t N, size_t Snew, size_t S,
                               size_t H) {
    MARK_BODY(double)
}

void oracle_mark_redundant_f32(const unsigned char *new_states, float *new_lpj,
                               const unsigned char *old_states, size_t N, size_t Snew, size_t S,
                               size_t H) {
    MARK_BODY(float)
}
