"""Runs the self-checking tcgen05 TF32 probe (qb200_bench_tcgen05_tf32): max abs error must be 0."""
import sys
sys.path.insert(0, ".")
import qrochet_b200 as qb
c = qb.Context(0)
err, tf128, tf256 = c.tcgen05_tf32_probe()
print(f"tcgen05.mma kind::tf32 probe: max |D - expected| = {err}, issue-bound {tf128:.1f} TFLOP/s (M=128, N=128), "
      f"{tf256:.1f} TFLOP/s (M=128, N=256)")
if "--i8" in sys.argv:
    err, t128, t256 = c.tcgen05_i8_probe()
    print(f"tcgen05.mma kind::i8 probe: max |D - expected| = {err}, issue-bound {t128:.1f} TOP/s (M=128, N=128), "
          f"{t256:.1f} TOP/s (M=128, N=256)")
