"""Dumps the coefficient tables of glibc 2.39's single-precision powf / logf (x86-64 FMA variants) out of the
libm.so.6 of this image, as the C initialisers of crazyara_b200/csrc/glibc_flt32.cuh.

The algorithms are glibc's sysdeps/ieee754/flt-32/{e_powf.c,e_logf.c,e_exp2f_data.c,e_powf_log2_data.c,e_logf_data.c}
(Szabolcs Nagy's ARM optimized-routines); the addresses below were read off the disassembly of __powf_fma / __logf_fma
(objdump -d libm.so.6: the ifunc resolvers at `powf` / `logf` return them on FMA+AVX2 machines).  The port itself is
pinned against the live libm by tests/test_glibc_flt32.py, so a wrong table cannot survive.
"""
import struct
import sys

LIBM = "/lib/x86_64-linux-gnu/libm.so.6"
# virtual address == file offset for .rodata of this libm (readelf -S: addr 0x8f000, off 0x8f000)
POWF_LOG2_TAB, POWF_LOG2_POLY = 0xB7F80, 0xB8080   # 16 x {invc, logc}; 5 doubles
EXP2F_TAB, EXP2F_SHIFT_SCALED, EXP2F_POLY = 0xB7BE0, 0xB7CE0, 0xB7CE8   # 32 x u64; 1; 3 doubles
LOGF_TAB, LOGF_LN2, LOGF_POLY = 0xB7D40, 0xB7E40, 0xB7E48   # 16 x {invc, logc}; 1; 3 doubles


def main():
    data = open(LIBM, "rb").read()
    d = lambda off, n: struct.unpack_from("<%dd" % n, data, off)
    q = lambda off, n: struct.unpack_from("<%dQ" % n, data, off)
    out = []
    fmt = lambda xs: ", ".join(float(x).hex() for x in xs)
    out.append("#define ARA_POWF_LOG2_TAB {%s}" % fmt(d(POWF_LOG2_TAB, 32)))
    out.append("#define ARA_POWF_LOG2_POLY {%s}" % fmt(d(POWF_LOG2_POLY, 5)))
    out.append("#define ARA_EXP2F_TAB {%s}" % ", ".join("0x%016xULL" % x for x in q(EXP2F_TAB, 32)))
    out.append("#define ARA_EXP2F_SHIFT_SCALED %s" % fmt(d(EXP2F_SHIFT_SCALED, 1)))
    out.append("#define ARA_EXP2F_POLY {%s}" % fmt(d(EXP2F_POLY, 3)))
    out.append("#define ARA_LOGF_TAB {%s}" % fmt(d(LOGF_TAB, 32)))
    out.append("#define ARA_LOGF_LN2 %s" % fmt(d(LOGF_LN2, 1)))
    out.append("#define ARA_LOGF_POLY {%s}" % fmt(d(LOGF_POLY, 3)))
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
