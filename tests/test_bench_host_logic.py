"""CPU checks of bench.py's host-side helpers (no GPU): the config #1 scalar generator against the oracle's field
arithmetic, the sharding-independent scalar blocks, and the exact NTT multiply count used for `int_pipe`."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_geometric_scalars_follow_pippenger_bench(orc):
    """bb/plonk/pippenger_bench/main.cpp:52-58: accumulator = element; accumulator *= element per scalar"""
    from oracle import pyoracle as po
    b = _bench()
    sc = b.geometric_scalars(64, seed=7)
    ints = orc.from_mont_ints(po.FR, sc)
    e = ints[1] * pow(ints[0], -1, b.FR_MODULUS) % b.FR_MODULUS
    assert ints[0] == e * e % b.FR_MODULUS
    for i in range(1, 64):
        assert ints[i] == ints[i - 1] * e % b.FR_MODULUS
    # Montgomery form: one device/oracle multiply by the stored limbs gives the same product as the integers
    prod = orc.from_mont_ints(po.FR, orc.field_op(po.FR, po.OP_MUL, sc[:1], sc[1:2]))[0]
    assert prod == ints[0] * ints[1] % b.FR_MODULUS


def test_block_scalars_do_not_depend_on_sharding():
    import inputs
    b = _bench()
    blk = 1 << b.BLOCK_LOG
    whole = b.block_scalars(inputs, blk - 5, 10)
    assert np.array_equal(whole[:5], inputs.fr_elements(1000, blk)[-5:])
    assert np.array_equal(whole[5:], inputs.fr_elements(1001, blk)[:5])
    assert np.array_equal(b.block_scalars(inputs, blk, 5), whole[5:])


def test_ntt_multiply_count_model():
    b = _bench()
    # 2^22 = 8 + 7 + 7 bits, E = 4: (22 - 3 * 1.5) / 2 + 2 inter-pass twiddles
    assert abs(b.ntt_muls_per_element(22) - 10.75) < 1e-9
    # 2^16 = 8 + 8 bits, E = 2: (16 - 2 * 1) / 2 + 1
    assert abs(b.ntt_muls_per_element(16) - 8.0) < 1e-9
    assert b.ntt_passes(24) == 3 and b.ntt_passes(25) == 4
