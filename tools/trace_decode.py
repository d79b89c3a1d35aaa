"""All-CTA timeline of the generation-4 persistent decode kernel (BGPT_MEGA_PROF=1).

Every CTA stamps clock64 at the same events; two (globaltimer edge, clock64) pairs per CTA put the
per-SM clocks on one nanosecond axis.  For each event the table gives when the FIRST, the MEDIAN
and the LAST CTA reached it, relative to the moment the last CTA published the previous layer's
fc2 output (mean over layers 1..L-1).  The critical path of a layer is the chain of "last" rows.
  BGPT_MEGA_PROF=1 python tools/trace_decode.py --n-past 511
"""
import argparse
import os
import sys

os.environ.setdefault("BGPT_MEGA_PROF", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import importlib  # noqa: E402
import numpy as np  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ftype", default="q4_0")
ap.add_argument("--n-past", type=int, default=511)
ap.add_argument("--warm", type=int, default=4)
ap.add_argument("--dump", default="")
a = ap.parse_args()
capi = importlib.import_module("biogpt_cpp_b200.capi")
M = capi.Model.load(bench.model_path(a.ftype))
assert M.decode_generation == 4
M.decode_greedy(2, a.n_past, a.warm)
ids, ms = M.decode_greedy(2, a.n_past + a.warm, 1)
st, cal = M.read_trace()
M.close()
if a.dump:
    np.savez_compressed(a.dump, st=st, cal=cal)
nC, L1 = st.shape[0], st.shape[1]
L = L1 - 1
gbase = int(cal[:, 0].min())                               # epoch nanoseconds exceed float64's integer range: subtract first
g0, c0, g1, c1 = ((cal[:, i] - (gbase if i in (0, 2) else 0)).astype(np.float64) for i in range(4))
f = (c1 - c0) / (g1 - g0)                                  # cycles per ns of each SM
print(f"{a.ftype} n_past={a.n_past + a.warm}: kernel {ms * 1e3:.1f} us; SM clock {f.mean():.4f} GHz (min {f.min():.4f} max {f.max():.4f}); "
      f"launch skew of CTA starts {g0.max() - g0.min():.0f} ns")
valid = st > 0
T = g0[:, None, None, None] + (st - cal[:, 1][:, None, None, None]).astype(np.float64) / f[:, None, None, None]   # ns since the first CTA's start
T[~valid] = np.nan
ev = [((0, 0), "P1 tile start"), ((0, 3), "P1 x polled (warp 0)"), ((0, 4), "P1 all polled (sync 1)"), ((0, 6), "P1 LN before sync 2"), ((0, 7), "P1 LN after sync 2"), ((0, 8), "P1 LN scale known"), ((0, 5), "P1 LN + quantise done"),
      ((0, 1), "P1 record+weights ready"), ((0, 9), "P1 act loaded"), ((0, 10), "P1 dots done"), ((0, 2), "P1 q,k,v published"),
      ((1, 3), "att q,k,v polled"), ((1, 4), "att softmax done"), ((1, 2), "att block published"),
      ((2, 0), "P3 tile start"), ((2, 3), "P3 record polled (warp 0)"), ((2, 5), "P3 weights waited"), ((2, 1), "P3 record ready"), ((2, 9), "P3 act loaded"), ((2, 10), "P3 dots done"), ((2, 2), "P3 x1 published"),
      ((3, 0), "P4 tile start"), ((3, 3), "P4 x1 polled (warp 0)"), ((3, 4), "P4 all polled (sync 1)"), ((3, 6), "P4 LN before sync 2"), ((3, 7), "P4 LN after sync 2"), ((3, 8), "P4 LN scale known"), ((3, 5), "P4 LN + quantise done"),
      ((3, 1), "P4 record+weights ready"), ((3, 9), "P4 act loaded"), ((3, 10), "P4 dots done"), ((3, 2), "P4 block published"),
      ((4, 0), "P5 tile start"), ((4, 3), "P5 record polled (warp 0)"), ((4, 5), "P5 weights waited"), ((4, 1), "P5 record ready"),
      ((4, 4), "P5 phase A done"), ((4, 10), "P5 dots done"), ((4, 2), "P5 x published")]
ref = np.nanmax(T[:, :L, 4, 2], axis=0)                   # [L]: last CTA's fc2 publish of each layer
rows = []
for (ph, k), nme in ev:
    d = T[:, 1:L, ph, k] - ref[None, :L - 1]               # layers 1..L-1 relative to the previous layer's end
    if np.all(np.isnan(d)):
        continue
    with np.errstate(all="ignore"):
        first = np.nanmean(np.nanmin(d, axis=0)); med = np.nanmean(np.nanmedian(d, axis=0)); last = np.nanmean(np.nanmax(d, axis=0))
        who = int(np.nanargmax(np.nanmean(d, axis=1)))
    rows.append((last, first, med, nme, who, int(np.sum(~np.isnan(d[:, 0])))))
print(f"{'event':32s} {'first':>8s} {'median':>8s} {'last':>8s}  (ns after the previous layer's last fc2 publish; CTAs stamping; latest CTA)")
for last, first, med, nme, who, n in rows:
    print(f"{nme:32s} {first:8.0f} {med:8.0f} {last:8.0f}  {n:4d} {who:4d}")
print(f"layer period {np.diff(ref).mean():.0f} ns; lm_head + tail {np.nanmax(T[:, L, 0, 2]) - ref[L - 1]:.0f} ns; "
      f"first layer starts {np.nanmin(T[:, 0, 0, 0]) - g0.min():.0f} ns after launch, its fc2 done at {ref[0] - g0.min():.0f} ns")
# per-CTA busy intervals on the chain: durations between consecutive stamps of the same CTA (cycles)
C = st.astype(np.float64); C[~valid] = np.nan
def dur(a_, b_):
    return np.nanmean(C[:, 1:L, b_[0], b_[1]] - C[:, 1:L, a_[0], a_[1]], axis=1)
with np.errstate(all="ignore"):
    for nme, a_, b_ in [("P1 LN: sync1 -> before sync2", (0, 4), (0, 6)), ("P1 LN: sync2 wait", (0, 6), (0, 7)), ("P1 LN: sync2 -> scale", (0, 7), (0, 8)), ("P1 LN: scale -> quantised", (0, 8), (0, 5)),
                        ("P1 LN: sync1 -> before sync2, warp 8", (0, 4), (0, 11)), ("P4 LN: sync1 -> before sync2, warp 8", (3, 4), (3, 11)),
                        ("P4 LN: sync1 -> before sync2", (3, 4), (3, 6)), ("P4 LN: sync2 wait", (3, 6), (3, 7)), ("P4 LN: sync2 -> scale", (3, 7), (3, 8)), ("P4 LN: scale -> quantised", (3, 8), (3, 5)),
                        ("P1 ready -> act loaded", (0, 1), (0, 9)), ("P1 act loaded -> dots", (0, 9), (0, 10)), ("P1 dots -> published", (0, 10), (0, 2)),
                        ("P4 ready -> act loaded", (3, 1), (3, 9)), ("P4 act loaded -> dots", (3, 9), (3, 10)), ("P4 dots -> published", (3, 10), (3, 2)),
                        ("P3 act loaded -> dots", (2, 9), (2, 10)), ("P3 dots -> published", (2, 10), (2, 2)), ("P5 phase A -> dots", (4, 4), (4, 10)), ("P5 dots -> published", (4, 10), (4, 2)),
                        ("P1 LN: sync1 -> LN+quant done", (0, 4), (0, 5)), ("P1 LN done -> ready (mbar+sync)", (0, 5), (0, 1)), ("P1 matmul+publish", (0, 1), (0, 2)),
                        ("att polled -> softmax done", (1, 3), (1, 4)), ("att softmax -> published", (1, 4), (1, 2)),
                        ("P3 ready -> published", (2, 1), (2, 2)), ("P4 sync1 -> LN done", (3, 4), (3, 5)), ("P4 ready -> published", (3, 1), (3, 2)),
                        ("P5 ready -> phase A done", (4, 1), (4, 4)), ("P5 phase A done -> published", (4, 4), (4, 2))]:
        d = dur(a_, b_)
        print(f"  {nme:34s} cycles: min {np.nanmin(d):7.0f} median {np.nanmedian(d):7.0f} max {np.nanmax(d):7.0f}")
    c = 40
    for nme, a_, b_ in [("P1 sync1 -> before sync2", (0, 4), (0, 6)), ("P1 sync2 wait", (0, 6), (0, 7)), ("P1 sync2 -> scale", (0, 7), (0, 8)), ("P1 scale -> quantised", (0, 8), (0, 5)),
                        ("P4 sync1 -> before sync2", (3, 4), (3, 6)), ("P4 sync2 wait", (3, 6), (3, 7)), ("P4 sync2 -> scale", (3, 7), (3, 8)), ("P4 scale -> quantised", (3, 8), (3, 5)),
                        ("P4 act loaded -> dots", (3, 9), (3, 10)), ("P4 dots -> published", (3, 10), (3, 2))]:
        print(f"  CTA {c} per layer, {nme:26s}", (C[c, :L, b_[0], b_[1]] - C[c, :L, a_[0], a_[1]]).astype(np.int64))
