from typing import Any, Dict


def get(d: Dict[Any, Any], *keys: Any):
    """`val1, val2 = get(my_dict, 'key1', 'key2')` (tvo/utils/_utils.py:8-20)."""
    return map(d.get, keys)
