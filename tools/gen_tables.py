#!/usr/bin/env python
"""Generate klara.jl_b200/csrc/klb_tables.h: the constant tables shared by the
device kernels and the CPU oracle (exp / log lookup tables and the 256-layer
ziggurat for the standard normal).

All values are computed with mpmath at 60 significant digits and stored as the
uint64 bit pattern of the correctly rounded double, so host and device see the
same bits.  Re-run with:   python tools/gen_tables.py
"""
import os
import struct
import mpmath as mp

mp.mp.dps = 60

def d2u(x):
    """bit pattern of float(x) (x is an mpf; float() rounds to nearest)."""
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]

# ----------------------------------------------------------------------------
# ziggurat, 256 layers (Marsaglia & Tsang 2000; Doornik 2005 layout)
# ----------------------------------------------------------------------------
NZ = 256
f = lambda x: mp.exp(-x * x / 2)
tail = lambda r: mp.sqrt(mp.pi / 2) * mp.erfc(r / mp.sqrt(2))    # int_r^inf f

def closure(r):
    v = r * f(r) + tail(r)
    x = r
    for _ in range(NZ - 2):                # x[2] .. x[255]
        x = mp.sqrt(-2 * mp.log(v / x + f(x)))
    return v / x + f(x) - 1               # top layer must reach f(0) = 1

r = mp.findroot(closure, mp.mpf("3.6541528853610088"))
v = r * f(r) + tail(r)
xs = [v / f(r), r]                         # x[0] (virtual base width), x[1] = r
for _ in range(NZ - 2):
    xs.append(mp.sqrt(-2 * mp.log(v / xs[-1] + f(xs[-1]))))
xs.append(mp.mpf(0))                       # x[256] = 0
assert len(xs) == NZ + 1
ZX = [d2u(x) for x in xs[:NZ]]
ZK = [int(mp.floor(mp.mpf(2) ** 52 * xs[i + 1] / xs[i])) for i in range(NZ)]
ZF = [d2u(f(x)) for x in xs]               # 257 entries, ZF[256] = 1.0

# ----------------------------------------------------------------------------
# exp: 2^(j/128) split as  scale-bits (exponent pre-biased by -j<<45) + tail
# ----------------------------------------------------------------------------
NE = 128
EXPT = []
for j in range(NE):
    h = mp.mpf(2) ** (mp.mpf(j) / NE)
    hd = float(h)
    tail_ = h / mp.mpf(hd) - 1             # relative tail: 2^(j/N) = hd*(1+tail)
    EXPT.append(d2u(tail_))
    EXPT.append((d2u(hd) - (j << 45)) & 0xFFFFFFFFFFFFFFFF)

# ----------------------------------------------------------------------------
# log: z in [OFF, 2*OFF) ~ [0.69, 1.38); 128 sub-intervals of 2^45 bit patterns.  OFF is chosen so
# that 1.0 is the midpoint (in bit-pattern space) of its interval: [1 - 2^-9, 1 + 2^-8).
# ----------------------------------------------------------------------------
NL = 128
OFF = 0x3FE6100000000000
LOGT = []
for i in range(NL):
    lo_bits = OFF + (i << 45)
    hi_bits = OFF + ((i + 1) << 45)
    lo = struct.unpack("<d", struct.pack("<Q", lo_bits))[0]
    hi = struct.unpack("<d", struct.pack("<Q", hi_bits))[0]
    if lo <= 1.0 < hi:
        invc = 1.0                          # log(1) must be exactly 0
    else:
        c = (mp.mpf(lo) + mp.mpf(hi)) / 2
        # keep invc short (24 significant bits) so that z*invc-1 is cheap to reason about
        invc = float(mp.mpf(1) / c)
        invc = struct.unpack("<d", struct.pack("<Q", (d2u(invc) + (1 << 27)) & ~((1 << 28) - 1)))[0]
    logc = -mp.log(mp.mpf(invc))
    LOGT.append(d2u(invc))
    LOGT.append(d2u(logc))

def fmt(name, arr, per=4):
    out = ["#define %s_LEN %d" % (name, len(arr))]
    body = []
    for i in range(0, len(arr), per):
        body.append("  " + ", ".join("0x%016xULL" % a for a in arr[i:i + per]))
    return out, body

# single blob: [ZF 258][EXPT 256][LOGT 256][ZXK interleaved 512]
# The {x, k} pairs come LAST so that a kernel can append the 256 sign-flipped pairs {-x, k} right behind them in its
# shared-memory copy and index the 512 entries with (w & 511): bit 8 of the word is the sign of the draw.
blob = []
off_zf = len(blob)
blob += ZF + [0]
off_exp = len(blob)
blob += EXPT
off_log = len(blob)
blob += LOGT
off_zxk = len(blob)
assert off_zxk % 2 == 0
for i in range(NZ):
    # k[i] is stored with the exponent bits of 1.0 on top, so that the acceptance test compares the bits of
    # t = 1.m directly (the kernels test the high words only and need no extra shift)
    assert 0 <= ZK[i] < (1 << 52)
    blob += [ZX[i], ZK[i] | (0x3FF << 52)]

consts = {
    "KLB_ZIG_R": d2u(r),
    "KLB_ZIG_RINV": d2u(1 / r),
    "KLB_INVLN2N": d2u(NE / mp.log(2)),
    "KLB_NEGLN2HIN": None, "KLB_NEGLN2LON": None,
    "KLB_LN2HI": None, "KLB_LN2LO": None,
    "KLB_LOG_OFF": OFF,
}
# -ln2/N split: hi has 32 trailing zero bits so kd*hi is exact for |kd| < 2^31
ln2n = mp.log(2) / NE
hi = struct.unpack("<d", struct.pack("<Q", d2u(ln2n) & ~((1 << 32) - 1)))[0]
consts["KLB_NEGLN2HIN"] = d2u(-mp.mpf(hi))
consts["KLB_NEGLN2LON"] = d2u(-(ln2n - mp.mpf(hi)))
ln2 = mp.log(2)
hi2 = struct.unpack("<d", struct.pack("<Q", d2u(ln2) & ~((1 << 32) - 1)))[0]
consts["KLB_LN2HI"] = d2u(mp.mpf(hi2))
consts["KLB_LN2LO"] = d2u(ln2 - mp.mpf(hi2))

here = os.path.dirname(os.path.abspath(__file__))
dst = os.path.join(here, "..", "klara.jl_b200", "csrc", "klb_tables.h")
with open(dst, "w") as fh:
    fh.write("/* GENERATED by tools/gen_tables.py -- do not edit.\n"
             " * Bit patterns of correctly rounded doubles (mpmath, 60 digits).\n"
             " * Layout of KLB_TAB (uint64 units):\n"
             " *   [%d,%d)        ziggurat f[i] = exp(-x[i]^2/2), i = 0..256 (+1 pad)\n"
             " *   [%d,%d)      exp table {tail bits, scale bits} x 128\n"
             " *   [%d,%d)      log table {invc bits, logc bits} x 128\n"
             " *   [%d,%d)     ziggurat {x[i] bits, k[i] | 0x3ff<<52} pairs, i = 0..255 (last: see tools/gen_tables.py)\n"
             " */\n" % (off_zf, off_exp, off_exp, off_log, off_log, off_zxk, off_zxk, len(blob)))
    fh.write("#ifndef KLB_TABLES_H\n#define KLB_TABLES_H\n#include <stdint.h>\n")
    fh.write("#define KLB_TAB_ZXK %d\n#define KLB_TAB_ZF %d\n#define KLB_TAB_EXP %d\n#define KLB_TAB_LOG %d\n"
             % (off_zxk, off_zf, off_exp, off_log))
    fh.write("#define KLB_TAB_LEN %d\n" % len(blob))
    for k, val in consts.items():
        fh.write("#define %s_BITS 0x%016xULL\n" % (k, val))
    fh.write("#ifndef KLB_TAB_QUAL\n#define KLB_TAB_QUAL static const\n#endif\n")
    fh.write("KLB_TAB_QUAL uint64_t KLB_TAB[KLB_TAB_LEN] = {\n")
    _, body = fmt("KLB_TAB", blob)
    fh.write(",\n".join(body))
    fh.write("\n};\n#endif /* KLB_TABLES_H */\n")
print("r =", mp.nstr(r, 20), " v =", mp.nstr(v, 20), " x[255] =", mp.nstr(xs[255], 12))
print("wrote", os.path.normpath(dst), len(blob), "u64")
