import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import __graft_entry__ as ge
pkg = ge.load_package()
n, c, h, w = 64, 3, 4000, 6000
dev = torch.device("cuda", 0)
host = torch.empty((n, c, h, w), dtype=torch.int16).pin_memory()
d = torch.empty((8, c, h, w), dtype=torch.int16, device=dev)
for k in range(0, n, 8):
    pkg.dev.synth_frames(d.data_ptr(), 8, c, h, w, k, 0, h, 0, 0x5349524C, torch.cuda.current_stream().cuda_stream)
    host[k:k+8].copy_(d)
torch.cuda.synchronize()
frames = host.numpy().view(np.uint16)
seq = pkg.Sequence(frames)
for it in range(3):
    t0 = time.perf_counter()
    res = pkg.stack_mean_with_rejection(seq, pkg.SIGMA, (4.0, 3.0), pkg.MULTIPLICATIVE)
    t1 = time.perf_counter()
    print("call %d: %.2f ms; h2d %.2f dev %.2f d2h %.2f total %.2f" % (it, (t1 - t0) * 1e3, res.t_h2d_ms, res.t_device_ms, res.t_d2h_ms, res.t_total_ms), file=sys.stderr)
    del res
