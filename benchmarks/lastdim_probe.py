import sys, json, torch
sys.path.insert(0, "/root/repo")
from qsparse_b200 import ops
dev = torch.device("cuda:0")
flush = torch.zeros(128 * 1024 * 1024, device=dev); flush_rd = torch.zeros(96 * 1024 * 1024, device=dev)
def timed(fn, iters=8):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.add_(1.0); flush_rd.max()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for rows, C in ((32768, 4096), (131072, 1024), (262144, 128), (1 << 20, 64), (1 << 22, 10)):
    x = torch.randn(rows, C, device=dev); y = torch.empty_like(x); n = x.numel()
    dec = torch.full((C,), 5.0, device=dev); lay = (rows, C, 1)
    mask = (torch.arange(C, device=dev) % 4 == 0)
    lines = torch.tensor([[-0.5, 0.5]] * C, device=dev)
    r = dict(shape=[rows, C])
    t = timed(lambda: ops.fq_pow2_fwd(x, dec, lay, out=y)); r["fq_pow2_ch"] = (round(t, 1), round(8 * n / t / 1e3 / 6457.4, 3))
    t = timed(lambda: ops.fq_line_fwd(x, lines, 8, True, lay, out=y)); r["fq_line_ch"] = (round(t, 1), round(8 * n / t / 1e3 / 6457.4, 3))
    t = timed(lambda: ops.fq_pow2_fwd(x, dec[:1], lay, mask=mask, out=y)); r["fq_pow2_mask75"] = (round(t, 1), round(8 * n / t / 1e3 / 6457.4, 3))
    t = timed(lambda: ops.reduce_stats(x, lay, abssum=True, absmax=True)); r["reduce_ch"] = (round(t, 1), round(4 * n / t / 1e3 / 6457.4, 3))
    t = timed(lambda: ops.reduce_stats(x, lay, minmax=True)); r["reduce_minmax_ch"] = (round(t, 1), round(4 * n / t / 1e3 / 6457.4, 3))
    t = timed(lambda: ops.ste_bwd(x, dec, True, 8, 0, lay, clamp_in_place=False, want_gx=True)); r["ste_bwd_ch"] = (round(t, 1), round(8 * n / t / 1e3 / 6457.4, 3))
    ops.set_tuning(15, 0)
    t = timed(lambda: ops.fq_pow2_fwd(x, dec, lay, out=y)); r["fq_pow2_ch_walk"] = (round(t, 1), round(8 * n / t / 1e3 / 6457.4, 3))
    t = timed(lambda: ops.fq_line_fwd(x, lines, 8, True, lay, out=y)); r["fq_line_ch_walk"] = (round(t, 1), round(8 * n / t / 1e3 / 6457.4, 3))
    ops.set_tuning(15, 1)
    print(json.dumps(r), flush=True)
    del x, y
