#!/usr/bin/env python
"""Generator for the register-resident 254-bit Montgomery multiply / square used by every kernel.

Emits aztec-2.0_b200/csrc/mont_asm.inc: one inline-PTX block per (field, op).  8 x 32-bit limbs,
R = 2^256, coarse output in [0, 2p) with no final subtraction -- the same integer
(a*b + m*p) / 2^256 the reference computes (bb/ecc/fields/field_impl_generic.hpp:392-499), so raw
limbs match barretenberg bit for bit.

Scheme (derived for Blackwell's IMAD.WIDE pipe): operand-scanning CIOS, one 32-bit limb of b per
step, with TWO accumulators so every multiply-add is a `mad.lo.cc / madc.hi.cc` pair on an aligned
64-bit lane that ptxas fuses into a single IMAD.WIDE.U32(.X):
    P : lanes at limb positions (0,1) (2,3) (4,5) (6,7)   <- products a_i*b_j, m*p_i with i even
    Q : lanes at limb positions (1,2) (3,4) (5,6) (7,8)   <- ... with i odd
After the reduction step limb 0 of the running value is zero; dividing by 2^32 swaps the roles of
P and Q (position parity flips).  The one limb that falls out of alignment (old P[1]) is added to
new P[0] and its carry is fed straight into the first Q-lane of the next step's carry chain.

`python gen_mont.py --selftest` runs the very same instruction list through a Python emulation of
the PTX carry semantics against big-integer arithmetic (no GPU needed).
"""
import os
import random
import sys

FIELDS = {
    "fq": dict(p=0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47),
    "fr": dict(p=0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001),
}
N = 8
MASK = 0xFFFFFFFF


def limbs(v):
    return [(v >> (32 * i)) & MASK for i in range(N)]


class Prog:
    """Tiny IR: list of (op, dst, a, b, c). Operands: register names (str) or ints (immediates)."""

    def __init__(self):
        self.ins = []
        self.ntmp = 0

    def tmp(self):
        self.ntmp += 1
        return "t%d" % (self.ntmp - 1)

    def emit(self, op, dst, a, b=None, c=None):
        self.ins.append((op, dst, a, b, c))


def gen_mul(p, sqr=False):
    """Returns Prog computing r[0..7] = a*b*R^-1 (coarse). Inputs a0..a7, b0..b7 (b=a when sqr)."""
    pl = limbs(p)
    ninv = (-pow(p, -1, 1 << 32)) & MASK
    pr = Prog()
    A = ["a%d" % i for i in range(N)]
    B = A if sqr else ["b%d" % i for i in range(N)]
    P = [pr.tmp() for _ in range(N)]
    Q = [pr.tmp() for _ in range(N)]
    L = None
    for j in range(N):
        bj = B[j]
        if j == 0:
            for i in range(0, N, 2):
                pr.emit("mul.lo", P[i], A[i], bj)
                pr.emit("mul.hi", P[i + 1], A[i], bj)
            for i in range(1, N, 2):
                pr.emit("mul.lo", Q[i - 1], A[i], bj)
                pr.emit("mul.hi", Q[i], A[i], bj)
        else:
            # stray limb (position 0) into P[0]; carry rides into the Q chain which starts at position 1
            pr.emit("add.cc", P[0], P[0], L)
            for i in range(1, N, 2):
                lo_c = Q[i - 1] if Q[i - 1] is not None else 0
                hi_c = Q[i] if Q[i] is not None else 0
                if Q[i - 1] is None:
                    Q[i - 1] = pr.tmp()
                if Q[i] is None:
                    Q[i] = pr.tmp()
                pr.emit("madc.lo.cc", Q[i - 1], A[i], bj, lo_c)
                # top lane cannot carry out: the running value is < 2^288
                pr.emit("madc.hi.cc" if i != N - 1 else "madc.hi", Q[i], A[i], bj, hi_c)
            for i in range(0, N, 2):
                pr.emit("mad.lo.cc" if i == 0 else "madc.lo.cc", P[i], A[i], bj, P[i])
                pr.emit("madc.hi.cc", P[i + 1], A[i], bj, P[i + 1])
            pr.emit("addc", Q[N - 1], Q[N - 1], 0)
        # Montgomery step: m = P[0] * (-p^-1) mod 2^32 ; += m*p ; limb 0 becomes zero
        m = pr.tmp()
        pr.emit("mul.lo", m, P[0], ninv)
        for i in range(1, N, 2):
            pr.emit("mad.lo.cc" if i == 1 else "madc.lo.cc", Q[i - 1], m, pl[i], Q[i - 1])
            pr.emit("madc.hi.cc" if i != N - 1 else "madc.hi", Q[i], m, pl[i], Q[i])
        for i in range(0, N, 2):
            pr.emit("mad.lo.cc" if i == 0 else "madc.lo.cc", P[i], m, pl[i], P[i])
            pr.emit("madc.hi.cc", P[i + 1], m, pl[i], P[i + 1])
        pr.emit("addc", Q[N - 1], Q[N - 1], 0)
        # divide by 2^32: P <- Q ; Q <- P shifted down one lane ; stray limb = old P[1]
        L = P[1]
        P, Q = Q, P[2:] + [None, None]
    # result limb k = Q[k-1] (old P shifted) + P[k] with the stray limb L at position 0
    R = ["r%d" % i for i in range(N)]
    pr.emit("add.cc", R[0], P[0], L)
    for k in range(1, N):
        q = Q[k - 1]
        op = "addc.cc" if k != N - 1 else "addc"
        pr.emit(op, R[k], P[k], q if q is not None else 0)
    return pr


def gen_sqr(p):
    """r = a*a*R^-1 (coarse), the same integer (a*a + m*p) / 2^256 as gen_mul(p, sqr=True), with 100 instead of 128 wide
    multiply-adds: the off-diagonal products a_i*a_j (i < j) are formed once, doubled with funnel shifts, and the diagonal
    squares ride in on the wide multiply-adds that add them; the 512-bit square's low half is then reduced with the same
    two-accumulator steps as gen_mul and the high half is added at the end ((T_lo + m p) / R + T_hi).

    Phase A keeps TWO arrays of aligned 64-bit lanes, one for even and one for odd limb positions (products a_i*a_j land at
    position i + j).  Row j (multiplier a_j, i > j) is one carry chain per array over consecutive lanes; its carry-out
    lands in the low limb of the next lane of the same array, which at that moment holds at most earlier carries (rows are
    taken in order and their top lanes never decrease), so one addc ends the chain."""
    pl = limbs(p)
    ninv = (-pow(p, -1, 1 << 32)) & MASK
    pr = Prog()
    A = ["a%d" % i for i in range(N)]
    lo, hi = {}, {}  # lane at limb position pos -> registers of limbs pos, pos + 1 (missing = zero)

    def reg(d, pos):
        if d.get(pos) is None:
            d[pos] = pr.tmp()
            return d[pos], 0
        return d[pos], d[pos]

    for j in range(N - 1):
        for par in (1, 0):  # the chain of odd-position lanes, then the even one (they are independent numbers)
            idxs = [i for i in range(j + 1, N) if (i + j) % 2 == par]
            if not idxs:
                continue
            for n_, i in enumerate(idxs):
                pos = i + j
                dl, cl = reg(lo, pos)
                dh, ch = reg(hi, pos)
                pr.emit("mad.lo.cc" if n_ == 0 else "madc.lo.cc", dl, A[i], A[j], cl)
                pr.emit("madc.hi.cc", dh, A[i], A[j], ch)
            top = idxs[-1] + j + 2
            dl, cl = reg(lo, top)
            pr.emit("addc", dl, cl, 0)
    # U = E + (O << 32): limb k gets the low limb of lane k and the high limb of lane k - 1 (one of each array)
    u = [None] * (2 * N)
    first = True
    for k in range(1, 2 * N):
        x = lo.get(k)
        y = hi.get(k - 1)
        if x is None and y is None:
            # nothing lands here (only possible above the top carry): keep the chain's carry
            u[k] = pr.tmp()
            pr.emit("add.cc" if first else ("addc.cc" if k != 2 * N - 1 else "addc"), u[k], 0, 0)
        else:
            u[k] = pr.tmp()
            pr.emit("add.cc" if first else ("addc.cc" if k != 2 * N - 1 else "addc"), u[k], x if x is not None else 0, y if y is not None else 0)
        first = False
    # T = 2 U + sum_i a_i^2 2^(64 i): the doubling is a funnel shift per limb, the squares come in on wide multiply-adds
    t = [None] * (2 * N)
    for k in range(1, 2 * N):
        t[k] = pr.tmp()
        if k == 1:
            pr.emit("add", t[k], u[k], u[k])
        else:
            pr.emit("shf", t[k], u[k - 1], u[k], 1)
    T = [pr.tmp() for _ in range(2 * N)]
    for i in range(N):
        pr.emit("mad.lo.cc" if i == 0 else "madc.lo.cc", T[2 * i], A[i], A[i], t[2 * i] if t[2 * i] is not None else 0)
        pr.emit("madc.hi.cc" if i != N - 1 else "madc.hi", T[2 * i + 1], A[i], A[i], t[2 * i + 1])
    # Montgomery-reduce the low half with gen_mul's steps (no product rows)
    P = T[:N]
    Q = [None] * N
    L = None
    for j in range(N):
        carry = False
        if j > 0:
            pr.emit("add.cc", P[0], P[0], L)
            carry = True
        m = pr.tmp()
        pr.emit("mul.lo", m, P[0], ninv)
        for n_, i in enumerate(range(1, N, 2)):
            lo_c = Q[i - 1] if Q[i - 1] is not None else 0
            hi_c = Q[i] if Q[i] is not None else 0
            if Q[i - 1] is None:
                Q[i - 1] = pr.tmp()
            if Q[i] is None:
                Q[i] = pr.tmp()
            pr.emit("madc.lo.cc" if (carry or n_ > 0) else "mad.lo.cc", Q[i - 1], m, pl[i], lo_c)
            pr.emit("madc.hi.cc" if i != N - 1 else "madc.hi", Q[i], m, pl[i], hi_c)
        for i in range(0, N, 2):
            pr.emit("mad.lo.cc" if i == 0 else "madc.lo.cc", P[i], m, pl[i], P[i])
            pr.emit("madc.hi.cc", P[i + 1], m, pl[i], P[i + 1])
        pr.emit("addc", Q[N - 1], Q[N - 1], 0)
        L = P[1]
        P, Q = Q, P[2:] + [None, None]
    # (T_lo + m p) / R = P + (Q shifted, stray limb at 0); then + T_hi
    S = [pr.tmp() for _ in range(N)]
    pr.emit("add.cc", S[0], P[0], L)
    for k in range(1, N):
        q = Q[k - 1]
        pr.emit("addc.cc" if k != N - 1 else "addc", S[k], P[k], q if q is not None else 0)
    R = ["r%d" % i for i in range(N)]
    for k in range(N):
        pr.emit("add.cc" if k == 0 else ("addc.cc" if k != N - 1 else "addc"), R[k], S[k], T[N + k])
    return pr


def emulate(pr, env):
    """Run the IR with PTX carry-flag semantics. env: dict reg -> value."""
    cc = 0

    def val(x):
        return x if isinstance(x, int) else env[x]

    for op, dst, a, b, c in pr.ins:
        base = op.split(".")
        name = base[0]
        if name in ("mul",):
            prod = val(a) * val(b)
            env[dst] = (prod & MASK) if base[1] == "lo" else (prod >> 32) & MASK
        elif name in ("mad", "madc"):
            prod = val(a) * val(b)
            part = (prod & MASK) if base[1] == "lo" else (prod >> 32) & MASK
            s = part + val(c) + (cc if name == "madc" else 0)
            env[dst] = s & MASK
            if op.endswith(".cc"):
                cc = s >> 32
        elif name in ("add", "addc"):
            s = val(a) + val(b) + (cc if name == "addc" else 0)
            env[dst] = s & MASK
            if op.endswith(".cc"):
                cc = s >> 32
        elif name == "shf":
            # shf.l.wrap.b32 d, lo, hi, n: the upper word of (hi:lo) << n
            env[dst] = ((val(b) << val(c)) | (val(a) >> (32 - val(c)))) & MASK
        else:
            raise ValueError(op)
        assert cc in (0, 1)
    return env


def selftest():
    random.seed(1)
    for fname, f in FIELDS.items():
        p = f["p"]
        rinv = pow(1 << 256, -1, p)
        for sqr in (False, True):
          for pr, label in ((gen_mul(p, sqr), "sqr" if sqr else "mul"),) + (((gen_sqr(p), "sqr (dedicated, --dedicated-sqr)"),) if sqr else ()):
            for trial in range(6000 if sqr else 3000):
                hi = 2 * p
                a = random.choice([0, 1, p - 1, p, p + 1, 2 * p - 1, random.randrange(hi), random.randrange(hi), random.randrange(hi),
                                   (2 * p - 1 - random.randrange(1 << 40)), random.randrange(1 << random.randrange(1, 255)),
                                   sum(random.choice([0, MASK, 0x80000000, 1]) << (32 * i) for i in range(N)) % hi])
                b = a if sqr else random.choice([0, 1, p - 1, p, 2 * p - 1, random.randrange(hi), random.randrange(hi)])
                env = {}
                for i, v in enumerate(limbs(a)):
                    env["a%d" % i] = v
                for i, v in enumerate(limbs(b)):
                    env["b%d" % i] = v
                emulate(pr, env)
                r = sum(env["r%d" % i] << (32 * i) for i in range(N))
                m = (a * b * ((-pow(p, -1, 1 << 256)) % (1 << 256))) % (1 << 256)
                exact = (a * b + m * p) >> 256
                assert r == exact, (fname, sqr, hex(a), hex(b), hex(r), hex(exact))
                assert r < 2 * p and r % p == (a * b * rinv) % p
            nmad = sum(1 for i in pr.ins if i[0].startswith(("mad", "mul")))
            nwide = sum(1 for i in pr.ins if i[0].startswith(("mad.hi", "madc.hi", "mul.hi")))
            print("%s %s: ok  (%d instructions, %d mul/mad = %d wide pairs + %d single)" % (
                fname, label, len(pr.ins), nmad, nwide, nmad - 2 * nwide))


def to_ptx(pr, fname, opname, sqr):
    """One asm() statement. Inputs/outputs bound by position: r0..r7 = %0..%7, a = %8.., b = %16.."""
    bind = {}
    for i in range(N):
        bind["r%d" % i] = "%%%d" % i
        bind["a%d" % i] = "%%%d" % (8 + i)
        if not sqr:
            bind["b%d" % i] = "%%%d" % (16 + i)

    def o(x):
        if isinstance(x, int):
            return "0x%08x" % x
        return bind.get(x, x)

    lines = []
    tmps = sorted({x for ins in pr.ins for x in ins[1:] if isinstance(x, str) and x.startswith("t")},
                  key=lambda s: int(s[1:]))
    lines.append(".reg .u32 %s;" % ", ".join(tmps))
    for op, dst, a, b, c in pr.ins:
        name = op.split(".")[0]
        if name == "mul":
            lines.append("%s.u32 %s, %s, %s;" % (op, o(dst), o(a), o(b)))
        elif name in ("mad", "madc"):
            lines.append("%s.u32 %s, %s, %s, %s;" % (op, o(dst), o(a), o(b), o(c)))
        elif name == "shf":
            lines.append("shf.l.wrap.b32 %s, %s, %s, %d;" % (o(dst), o(a), o(b), c))
        else:
            lines.append("%s.u32 %s, %s, %s;" % (op, o(dst), o(a), o(b)))
    body = "\n".join('        "%s\\n\\t"' % ln for ln in lines)
    outs = ", ".join('"=r"(r[%d])' % i for i in range(N))
    ins = ", ".join('"r"(a[%d])' % i for i in range(N))
    if not sqr:
        ins += ", " + ", ".join('"r"(b[%d])' % i for i in range(N))
    sig = "const uint32_t (&a)[8]" if sqr else "const uint32_t (&a)[8], const uint32_t (&b)[8]"
    return (
        "// %s %s: %d PTX instructions\n"
        "__device__ __forceinline__ void %s_%s_ptx(uint32_t (&r)[8], %s)\n{\n"
        "    asm(\"{\\n\\t\"\n%s\n        \"}\"\n        : %s\n        : %s);\n}\n"
        % (fname, opname, len(pr.ins), fname, opname, sig, body, outs, ins))


def main():
    if "--selftest" in sys.argv:
        selftest()
        return
    out = ["// GENERATED by gen_mont.py -- do not edit. See gen_mont.py for the derivation.",
           "#pragma once", "#include <cstdint>", ""]
    for fname, f in FIELDS.items():
        out.append(to_ptx(gen_mul(f["p"], False), fname, "mul", False))
        # The dedicated squaring (gen_sqr: 100 instead of 128 wide multiply-adds, bit-identical results, the whole GPU suite
        # passes with it) is NOT what ships: measured on B200 the MSM's accumulate kernel went from 2.14 to 2.18 ms with it --
        # 7 % fewer IMAD.WIDE, but +110 IADD3 / +75 SHF per kernel and 34 instead of 12 spilled bytes at the 128-register cap
        # cost more than the multiplier slots they free.  `--dedicated-sqr` emits it for experiments.
        dedicated = "--dedicated-sqr" in sys.argv
        out.append(to_ptx(gen_sqr(f["p"]) if dedicated else gen_mul(f["p"], True), fname, "sqr", True))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mont_asm.inc")
    with open(path, "w") as fh:
        fh.write("\n".join(out))
    print("wrote", path)


if __name__ == "__main__":
    main()
