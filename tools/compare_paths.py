"""Full-size config-2 batch through the batch path (ilqr_fit on one handle) and through the streamer: where do the
results differ, if anywhere?   python tools/compare_paths.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ilqr_b200  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402

H, B = 200, 65536
with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) as s:
    x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
    s.upload_x0(x0, np.zeros((H, 2, B), order="F"))
    dx = torch.empty((B, 4, H + 1), dtype=torch.float64, device="cuda")
    s.download_device(_abi.X, dx.data_ptr())
    du = torch.zeros((B, 2, H), dtype=torch.float64, device="cuda")
    s.upload_device(dx.data_ptr(), du.data_ptr())
    s.fit(100, 1e-6)
    bx, bu = torch.empty_like(dx), torch.empty_like(du)
    s.download_device(_abi.X, bx.data_ptr()); s.download_device(_abi.U, bu.data_ptr())
    bi, bs, bc = s.download(_abi.ITERS), s.download(_abi.STATUS), s.download(_abi.PREV_COST)
outs = [torch.zeros_like(dx), torch.zeros_like(du), torch.zeros(B, dtype=torch.float64, device="cuda"),
        torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")]
if not os.environ.get("SKIP_STREAMER"):
    with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, 56832), B, ring=2) as st:
        st.wait(st.submit_ptrs(dx.data_ptr(), du.data_ptr(), *[t.data_ptr() for t in outs], device=True))
else:
    outs = [bx.clone(), bu.clone(), torch.from_numpy(bc).cuda(), torch.from_numpy(bi).cuda(), torch.from_numpy(bs).cuda()]
torch.cuda.synchronize()
sx, su, sc, si, ss = [t.cpu().numpy() for t in outs]
bx, bu = bx.cpu().numpy(), bu.cpu().numpy()
print("iters equal:", np.array_equal(si, bi), " status equal:", np.array_equal(ss, bs), " cost equal:", np.array_equal(sc, bc, equal_nan=True))
print("nan in batch x/u:", int(np.isnan(bx).any(axis=(1, 2)).sum()), int(np.isnan(bu).any(axis=(1, 2)).sum()),
      " nan in streamer x/u:", int(np.isnan(sx).any(axis=(1, 2)).sum()), int(np.isnan(su).any(axis=(1, 2)).sum()))
dxm = ~np.all((sx == bx) | (np.isnan(sx) & np.isnan(bx)), axis=(1, 2))
dum = ~np.all((su == bu) | (np.isnan(su) & np.isnan(bu)), axis=(1, 2))
print("trajectories with differing x:", int(dxm.sum()), " u:", int(dum.sum()))
bad = np.nonzero(dxm | dum | (si != bi) | (ss != bs))[0]
for t in bad[:12]:
    print(" traj %d: iters %d/%d status %d/%d cost %.17g/%.17g max|dx| %.3e max|du| %.3e" %
          (t, bi[t], si[t], bs[t], ss[t], bc[t], sc[t], np.nanmax(np.abs(sx[t] - bx[t])), np.nanmax(np.abs(su[t] - bu[t]))))
print("status histogram batch:", dict(zip(*np.unique(bs, return_counts=True))))

# ---- the pool scheduler (8 handles, batches in flight concurrently) against the same reference
NF, STEPS = int(os.environ.get('NF', '8')), int(os.environ.get('STEPS', '16'))
bxd, bud = torch.from_numpy(bx).cuda(), torch.from_numpy(bu).cuda()
with ilqr_b200.SolverPool(ilqr_b200.two_link_problem(H, B), NF) as pool:
    pouts = [(torch.zeros_like(dx), torch.zeros_like(du), torch.zeros(B, dtype=torch.int32, device="cuda")) for _ in range(NF)]
    tickets = []
    for i in range(STEPS):
        if i >= NF:
            pool.wait(tickets[i - NF])
            torch.cuda.synchronize()
            ox, ou, oi = pouts[i % NF]
            ex = int((~((ox == bxd).all(dim=2).all(dim=1))).sum().item()); eu = int((~((ou == bud).all(dim=2).all(dim=1))).sum().item())
            ei = int((oi.cpu().numpy() != bi).sum())
            print("pool[NF=%d %s] step %d: trajectories differing in x %d, u %d, iters %d" % (NF, {k: v for k, v in os.environ.items() if k.startswith("ILQR_")}, i - NF, ex, eu, ei))
        ox, ou, oi = pouts[i % NF]
        tickets.append(pool.submit_ptrs(dx.data_ptr(), du.data_ptr(), None, 100, 1e-6, ox.data_ptr(), ou.data_ptr(), None, oi.data_ptr(), None, device=True))
    pool.wait_all()

# ---- two streamers on the same GPU at the same time (their round kernels overlap) against the same reference
if not os.environ.get("SKIP_TWO_STREAMERS"):
    import threading
    sts = [ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, 28416), B, ring=3) for _ in range(2)]
    souts = [[[torch.zeros_like(dx), torch.zeros_like(du), torch.zeros(B, dtype=torch.float64, device="cuda"),
               torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")] for _ in range(3)] for _ in range(2)]

    def run2(i):
        tk = [sts[i].submit_ptrs(dx.data_ptr(), du.data_ptr(), *[t.data_ptr() for t in souts[i][b]], device=True) for b in range(3)]
        for t in tk:
            sts[i].wait(t)

    th = [threading.Thread(target=run2, args=(i,)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    for i in range(2):
        for b in range(3):
            ox, ou, oc, oi, os_ = souts[i][b]
            ex = int((~((ox == bxd).all(dim=2).all(dim=1))).sum().item()); eu = int((~((ou == bud).all(dim=2).all(dim=1))).sum().item())
            print("streamer %d batch %d (two streamers concurrently): trajectories differing in x %d, u %d, iters %d" %
                  (i, b, ex, eu, int((oi.cpu().numpy() != bi).sum())))
    [s_.close() for s_ in sts]
