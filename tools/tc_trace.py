"""Timeline of CTA 0 of scan_tc_kernel (needs tools/libais_trace.so built with -DAIS_TC_TRACE)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ais_b200  # noqa
from ais_b200 import engine as E, synth, binding as B

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_docs = 128 * 148 * 60
idx = synth.generate_index(20000, vocab_size=500, seed=12)
eng = E.SearchEngine(device=0, max_batch=mb)
rows = eng.rows_tensor(n_docs)
rows.normal_()
import torch
sh_ptr = torch.zeros(501, dtype=torch.int64, device="cuda")
eng.load_bm25(sh_ptr, torch.zeros(0, dtype=torch.int32, device="cuda"), None, torch.zeros(500, dtype=torch.float64, device="cuda"),
              torch.ones(n_docs, dtype=torch.int64, device="cuda"), 1.0)
rng = np.random.default_rng(1)
vecs = rng.standard_normal((mb, 300)).astype(np.float32)
qs = [E.Query(v, np.array([0], np.int32), np.array([1.0])) for v in vecs]
maxes = torch.empty((mb, 2), dtype=torch.float64, device="cuda")
for _ in range(3):
    eng.stage_score(qs, maxes)
torch.cuda.synchronize(); eng.synchronize()
out = np.zeros((4, 256, 4), dtype=np.int64)
B.lib.ais_debug_tc_trace.restype = C.c_int
B.lib.ais_debug_tc_trace.argtypes = [C.c_void_p, C.c_void_p]
B.check(B.lib.ais_debug_tc_trace(eng._h, out.ctypes.data))
t0 = out[3, 0, 0]
print('t0', t0)
ep = out[0, :60, :3] - t0
print("epilogue per tile: [acc_full seen, released, stores done] and deltas; tile period")
for t in range(2, 16):
    print(t, ep[t], "wait->release %d, release->done %d, period %d" % (ep[t,1]-ep[t,0], ep[t,2]-ep[t,1], ep[t,0]-ep[t-1,0]))
mm = out[1, :256, :2] - t0
print("MMA per A stage: [a_full seen, issued]; period")
for i in range(40, 90):
    print(i, mm[i], "issue %d period %d" % (mm[i,1]-mm[i,0], mm[i,0]-mm[i-1,0]))
sp = out[2, :128, :4] - t0
print("split warp 4 per own iteration: [full_raw seen, computed, a_empty seen, arrived]")
for i in range(15, 45):
    print(i, sp[i], "compute %d, wait a_empty %d, st+arrive %d, period %d" % (sp[i,1]-sp[i,0], sp[i,2]-sp[i,1], sp[i,3]-sp[i,2], sp[i,0]-sp[i-1,0]))
pr = out[3, :256, 0] - t0
print("producer issue times (deltas):", np.diff(pr[40:80]))
