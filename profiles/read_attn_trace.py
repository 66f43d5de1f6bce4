"""Print the phase timeline written by B200_ATTN_TRACE=<file> (window_attention_tc, CTA 0, first 64 tiles; SM clock cycles).

roles: 0 = MMA issuer  (0 P0 ready -> issue PV0 | 1 S0 operands ready -> issue | 2 S0 issued | 3 P1 ready -> issue PV1 | 4 S1 operands ready | 5 S1 issued)
       1 = softmax warp of key half 0, 2 = of key half 1  (0 S ready | 1 pass 1 done | 2 maxima exchanged | 3 (unused) | 4 pass 2 done | 5 O ready (epilogue) | 6 epilogue done)"""
import sys

import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(64, 3, 8)
t0 = t[4, 0, 0]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4, 12)
for it in range(lo, hi):
    for role, name in enumerate(("mma", "sm0", "sm1")):
        ev = t[it, role]
        print(f"tile {it:2d} {name}: " + " ".join(f"{int(v - t0):7d}" if v else "      -" for v in ev[:7]))
per = np.diff(t[4:60, 0, 0]).astype(float)
print("tile period: median %.0f cycles (min %.0f, max %.0f)" % (np.median(per), per.min(), per.max()))
for hf in (0, 1):
    e = t[4:60, 1 + hf].astype(float)
    print(f"half {hf}: pass1 {np.median(e[:, 1] - e[:, 0]):.0f}, exchange {np.median(e[:, 2] - e[:, 1]):.0f}, "
          f"pass2 {np.median(e[:, 4] - e[:, 2]):.0f}, P done -> next S ready {np.median(e[1:, 0] - e[:-1, 4]):.0f}")
m = t[4:60, 0].astype(float)
print(f"mma: PV0 issue..S0 operands {np.median(m[:, 1] - m[:, 0]):.0f}, S0 issue {np.median(m[:, 2] - m[:, 1]):.0f}, wait P1 {np.median(m[:, 3] - m[:, 2]):.0f}, "
      f"PV1 issue..S1 operands {np.median(m[:, 4] - m[:, 3]):.0f}, S1 issue {np.median(m[:, 5] - m[:, 4]):.0f}, wait P0 of next {np.median(m[1:, 0] - m[:-1, 5]):.0f}")
