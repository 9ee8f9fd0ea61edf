"""Layer-by-layer comparison of the device detector with the oracle interpreter, plus per-step device times
(diagnostic script, run on a GPU box: `python tests/diag_dbnet.py [H W] [-r] [-v]`; not collected by pytest)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from oracle import dbnet_oracle as D
from vsr_b200.dbnet import TextDetector

MODEL = os.path.join("weights", "V5", "ch_det")
rng = np.random.default_rng(0)
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (720, 1280)
img = cv2.GaussianBlur(rng.integers(0, 255, (h, w, 3), dtype=np.uint8), (0, 0), 9)
cv2.putText(img, "The quick brown fox 0123", (int(w * .2), int(h * .9)), cv2.FONT_HERSHEY_SIMPLEX, 1.6, (255, 255, 255), 3, cv2.LINE_AA)
det = TextDetector(MODEL, "cuda:0")
got = det.probability_map(img)
vals = {}
want = D.forward(D.Graph(MODEL), D.preprocess(img), values=vals)[0, 0].numpy()
prog = next(iter(det._programs.values()))
prodkind = {n.out: n.kind + ":" + str(n.attrs.get("struct_name", ""))[-60:] for n in det._nodes if n.out is not None}
print("final: max diff", np.abs(got - want).max(), "nonfinite", (~np.isfinite(got)).sum())
bad = 0
tail_ids = set(sorted(prog.values)[-14:])
for vid in sorted(prog.values):
    t = prog.values[vid]
    if vid not in vals:
        continue
    ref = vals[vid][0].permute(1, 2, 0).numpy()
    dev = det._rt.download(t).astype(np.float32) / t.scale
    idx = t.perm if t.perm is not None else np.arange(t.c)
    dev = dev[:, :, idx]
    nf = int((~np.isfinite(dev)).sum())
    d = np.abs(np.nan_to_num(dev, posinf=1e9, neginf=-1e9) - ref)
    scale = np.abs(ref).max()
    flag = "  <<<<" if (nf or d.max() > 0.05 * max(scale, 1.0)) else ""
    if flag:
        bad += 1
    if "-r" in sys.argv:
        print(f"R id {vid:5d} c={t.c:4d} {t.h}x{t.w} scale={t.scale:g} ref|max|={scale:10.3f} ref_rms={np.sqrt((ref ** 2).mean()):10.4f} "
              f"rel_rms={np.sqrt((d ** 2).mean()) / (np.sqrt((ref ** 2).mean()) + 1e-12):.2e} by {prodkind.get(vid)}")
    if vid in tail_ids:
        y, x, c = np.unravel_index(np.argmax(d), d.shape)
        print(f"   tail id {vid}: scale={t.scale} argmax diff at (y={y}, x={x}, c={c}) ref={ref[y, x, c]:.4f} dev={dev[y, x, c]:.4f}; "
              f"count(diff>0.05*max)={(d > 0.05 * max(scale, 1.0)).sum()} rel_rms={np.sqrt((d ** 2).mean()) / (np.sqrt((ref ** 2).mean()) + 1e-12):.2e}")
    if flag or "-v" in sys.argv or vid in tail_ids:
        print(f"id {vid:5d} c={t.c:4d} {t.h}x{t.w} cp={t.cp} ref|max|={scale:9.3f} maxdiff={d.max():9.4f} mean={d.mean():.5f} nonfinite={nf}{flag}")
    if bad >= 40:
        break
# per-step device time (eager, one sync per step): where the frame time goes
import time
from collections import defaultdict
rt = det._rt
rt.L.vsr_rt_sync(rt.h)
acc = defaultdict(float)
rows = []
for st in prog.steps:
    t0 = time.perf_counter()
    for _ in range(3):
        st.run(rt)
    rt.L.vsr_rt_sync(rt.h)
    dt = (time.perf_counter() - t0) / 3 * 1e3
    name = type(st).__name__
    y = getattr(st, "y", None) or (st.args[1] if hasattr(st, "args") else None)
    desc = f"{name} -> {y.c}x{y.h}x{y.w}" if y is not None else name
    if name == "_Conv":
        desc += f" from {st.x.c}ch lid={st.lid}"
    acc[name] += dt
    rows.append((dt, desc))
print("per-kind ms:", {k: round(v, 3) for k, v in acc.items()}, "total", round(sum(acc.values()), 3))
for dt, desc in sorted(rows, reverse=True)[:25]:
    print(f"  {dt:8.3f} ms  {desc}")
t0 = time.perf_counter(); rt.preprocess(img, prog.inp, prog.inp.h, prog.inp.w); rt.L.vsr_rt_sync(rt.h); print("preprocess ms", (time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter(); rt.graph_launch(prog.graph); rt.L.vsr_rt_sync(rt.h); print("graph ms", (time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter(); rt.download(prog.out); print("download ms", (time.perf_counter() - t0) * 1e3)
