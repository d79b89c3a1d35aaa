"""All-CTA timeline of the generation-5 persistent decode kernel (BGPT_MEGA_PROF=1).

Every CTA stamps clock64 at the same events; two (globaltimer edge, clock64) pairs per CTA put the per-SM clocks on one
nanosecond axis.  For each event the table gives when the FIRST, the MEDIAN and the LAST CTA reached it, relative to the moment
the last CTA published the previous layer's fc2 output (mean over layers 1..L-1).  The critical path of a layer is the chain
of "last" rows.
  BGPT_MEGA_PROF=1 python tools/trace_decode5.py --n-past 511
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
assert M.decode_generation == 5, M.decode_generation
M.decode_greedy(2, a.n_past, a.warm)
ids, ms = M.decode_greedy(2, a.n_past + a.warm, 1)
st, cal = M.read_trace()
M.close()
if a.dump:
    np.savez_compressed(a.dump, st=st, cal=cal)
nC, L1 = st.shape[0], st.shape[1]
L = L1 - 1
gbase = int(cal[:, 0].min())
g0, c0, g1, c1 = ((cal[:, i] - (gbase if i in (0, 2) else 0)).astype(np.float64) for i in range(4))
f = (c1 - c0) / (g1 - g0)                                  # cycles per ns of each SM
print(f"{a.ftype} n_past={a.n_past + a.warm}: kernel {ms * 1e3:.1f} us; SM clock {f.mean():.4f} GHz; launch skew of CTA starts {g0.max() - g0.min():.0f} ns")
valid = st > 0
T = g0[:, None, None, None] + (st - cal[:, 1][:, None, None, None]).astype(np.float64) / f[:, None, None, None]
T[~valid] = np.nan
ev = [((0, 0), "P1 tile start"), ((0, 3), "P1 x polled (warp 0)"), ((0, 4), "P1 LayerNorm done"), ((0, 5), "P1 quantised"), ((0, 1), "P1 record+weights ready"),
      ((0, 10), "P1 dots done"), ((0, 2), "P1 q,k stored to the cluster"),
      ((1, 3), "att q,k arrived (cluster barrier)"), ((1, 5), "att scores arrived (cluster barrier)"), ((1, 4), "att softmax done"), ((1, 2), "att output published"),
      ((2, 0), "P3 tile start"), ((2, 3), "P3 att polled (warp 0)"), ((2, 5), "P3 quantised"), ((2, 1), "P3 record+weights ready"), ((2, 10), "P3 dots done"), ((2, 2), "P3 x1 published"),
      ((3, 0), "P4 tile start"), ((3, 3), "P4 x1 polled (warp 0)"), ((3, 4), "P4 LayerNorm done"), ((3, 5), "P4 quantised"), ((3, 1), "P4 record+weights ready"),
      ((3, 10), "P4 dots done"), ((3, 2), "P4 block published"),
      ((4, 0), "P5 tile start"), ((4, 3), "P5 blocks polled+scattered (warp 0)"), ((4, 1), "P5 record+weights ready"), ((4, 10), "P5 dots done"), ((4, 2), "P5 x published")]
ref = np.nanmax(T[:, :L, 4, 2], axis=0)                   # [L]: last CTA's fc2 publish of each layer
print(f"{'event':40s} {'first':>8s} {'median':>8s} {'last':>8s}  (ns after the previous layer's last fc2 publish; CTAs stamping; latest CTA)")
for (ph, k), nme in ev:
    d = T[:, 1:L, ph, k] - ref[None, :L - 1]
    if np.all(np.isnan(d)):
        continue
    with np.errstate(all="ignore"):
        first = np.nanmean(np.nanmin(d, axis=0)); med = np.nanmean(np.nanmedian(d, axis=0)); last = np.nanmean(np.nanmax(d, axis=0))
        who = int(np.nanargmax(np.nanmean(d, axis=1)))
    print(f"{nme:40s} {first:8.0f} {med:8.0f} {last:8.0f}  {int(np.sum(~np.isnan(d[:, 0]))):4d} {who:4d}")
print(f"layer period {np.diff(ref).mean():.0f} ns; lm_head + tail {np.nanmax(T[:, L, 0, 2]) - ref[L - 1]:.0f} ns; "
      f"first layer starts {np.nanmin(T[:, 0, 0, 0]) - g0.min():.0f} ns after launch, its fc2 done at {ref[0] - g0.min():.0f} ns")
C = st.astype(np.float64); C[~valid] = np.nan
def dur(a_, b_):
    return np.nanmean(C[:, 1:L, b_[0], b_[1]] - C[:, 1:L, a_[0], a_[1]], axis=1)
with np.errstate(all="ignore"):
    for nme, a_, b_ in [("P1 LN: polled -> before sync 1", (0, 3), (0, 7)), ("P1 LN: sync 1 -> mean known", (0, 7), (0, 8)), ("P1 LN: mean -> before sync 2", (0, 8), (0, 9)),
                        ("P1 LN: sync 2 -> variance known", (0, 9), (0, 11)), ("P1 LN: variance -> y", (0, 11), (0, 4)),
                        ("P1 quantised -> weights waited", (0, 5), (0, 6)), ("P1 weights waited -> ready", (0, 6), (0, 1)),
                        ("P3 quantised -> weights waited", (2, 5), (2, 6)), ("P3 weights waited -> ready", (2, 6), (2, 1)),
                        ("P1 polled -> LayerNorm done", (0, 3), (0, 4)), ("P1 LayerNorm -> quantised", (0, 4), (0, 5)), ("P1 quantised -> ready", (0, 5), (0, 1)),
                        ("P1 ready -> dots", (0, 1), (0, 10)), ("P1 dots -> stored", (0, 10), (0, 2)), ("P1 stored -> q,k arrived", (0, 2), (1, 3)),
                        ("att q,k -> scores arrived", (1, 3), (1, 5)), ("att scores -> softmax", (1, 5), (1, 4)), ("att softmax -> published", (1, 4), (1, 2)),
                        ("P3 polled -> quantised", (2, 3), (2, 5)), ("P3 quantised -> ready", (2, 5), (2, 1)), ("P3 ready -> dots", (2, 1), (2, 10)), ("P3 dots -> published", (2, 10), (2, 2)),
                        ("P4 polled -> LayerNorm done", (3, 3), (3, 4)), ("P4 LayerNorm -> quantised", (3, 4), (3, 5)), ("P4 quantised -> ready", (3, 5), (3, 1)),
                        ("P4 ready -> dots", (3, 1), (3, 10)), ("P4 dots -> published", (3, 10), (3, 2)),
                        ("P5 polled -> ready", (4, 3), (4, 1)), ("P5 ready -> dots", (4, 1), (4, 10)), ("P5 dots -> published", (4, 10), (4, 2))]:
        d = dur(a_, b_)
        print(f"  {nme:34s} cycles: min {np.nanmin(d):7.0f} median {np.nanmedian(d):7.0f} max {np.nanmax(d):7.0f}")
