"""Generates the fp64 lookup tables of eeyore_b200/csrc/common.cuh (correctly rounded, mpmath at 200 bits):
   exp_table.inc : 2^(j/2048), j = 0..2047                       (exp_core / sigmoid: degree-3 polynomial on |r| <= ln2/4096)
   log_table.inc : 1024 pairs (1/c_i rounded to double, -log of that double) over the window [sqrt(1/2), sqrt(2))
                   addressed by the high word like fdlibm's e_log.c normalisation (log_pos_normal: degree-4 polynomial).
   python tools/gen_math_tables.py"""
import struct
from pathlib import Path

import mpmath as mp

mp.mp.prec = 200
OUT = Path(__file__).resolve().parents[1] / "eeyore_b200" / "csrc"


def dbl(v):
    return float(mp.nstr(v, 40))


def fmt(vals, per=4):
    lines = []
    for i in range(0, len(vals), per):
        lines.append("  " + ", ".join(repr(v) for v in vals[i:i + per]) + ",")
    return "\n".join(lines) + "\n"


def main():
    e = [dbl(mp.power(2, mp.mpf(j) / 2048)) for j in range(2048)]
    (OUT / "exp_table.inc").write_text("// 2^(j/2048), j = 0..2047, correctly rounded (tools/gen_math_tables.py)\n" + fmt(e))

    # window position f = (hi + 0x95f64) & 0xfffff  <->  normalised m with hi word f + 0x3ff00000 - 0x95f64
    OFF = 0x3FF00000 - 0x95F64

    def m_of(f, lo):
        return struct.unpack("<d", struct.pack("<Q", ((f + OFF) << 32) | lo))[0]

    inv, nlog, rmax = [], [], 0.0
    for i in range(1024):
        lo_m, hi_m = m_of(i * 1024, 0), m_of(i * 1024 + 1023, 0xFFFFFFFF)
        if lo_m <= 1.0 <= hi_m:
            c = mp.mpf(1)                      # the bin that straddles 1: r = m - 1 exactly, no cancellation near x = 1
        else:
            c = (mp.mpf(lo_m) + mp.mpf(hi_m)) / 2
        ic = dbl(1 / c)
        inv.append(ic)
        nlog.append(dbl(-mp.log(mp.mpf(ic))) if ic != 1.0 else 0.0)
        for m in (lo_m, hi_m):
            rmax = max(rmax, abs(float(mp.mpf(m) * mp.mpf(ic) - 1)))
    print("log table: max |r| =", rmax, " r^5/5 =", rmax ** 5 / 5)
    (OUT / "log_table.inc").write_text(
        "// log_pos_normal tables over m in [sqrt(1/2), sqrt(2)): entry i = window position >> 10 (tools/gen_math_tables.py)\n"
        "#define EB_LOG_INV_TABLE \\\n" + fmt(inv).replace("\n", " \\\n") + "\n"
        "#define EB_LOG_NLOG_TABLE \\\n" + fmt(nlog).replace("\n", " \\\n") + "\n")


if __name__ == "__main__":
    main()
