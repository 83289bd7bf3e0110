import numpy as np


def unhex(s, width=4):
    """inverse of gen_golden.hx: 16 hex chars per u64 limb, limbs in memory order."""
    vals = [int(s[i:i + 16], 16) for i in range(0, len(s), 16)]
    a = np.array(vals, dtype=np.uint64)
    return a.reshape(-1, width) if width else a


def msm_scalars(case, checker):
    """Rebuild the scalar array of a golden MSM case (tests/golden/gen_golden.py)."""
    import inputs
    from oracle import pyoracle as po
    n, seed, kind = case["n"], case["seed"], case["kind"]
    if kind in ("uniform", "repeated_point7"):
        return inputs.fr_elements(seed, n)
    if kind == "coarse":
        return inputs.fr_elements(seed, n, coarse_fraction=0.5)
    if kind == "short":
        return checker.field_op(po.FR, po.OP_TO_MONT, po.ints_to_array(inputs.short_scalar_ints(seed, n)))
    if kind == "zeros":
        return np.zeros((n, 4), dtype=np.uint64)
    if kind == "ones":
        return checker.field_op(po.FR, po.OP_TO_MONT, po.ints_to_array([1] * n))
    raise ValueError(kind)
