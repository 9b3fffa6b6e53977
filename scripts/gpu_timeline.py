"""Per-CTA phase timeline of the TS-mode GEMM (AO_B200_TIMELINE=1).  Prints, per shape, the median/max over
CTAs of each phase (cycles from kernel entry): 1 prologue done, 2 pdl_wait returned, 3 first weights landed,
4 first MMA issued, 5 last MMA issued, 6 last accumulator ready, 7 epilogue done; plus the spread of entry times."""
import os
import sys

os.environ["AO_B200_TIMELINE"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
ops = torch.ops.ao_b200
g = 32


def mk(N, K):
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
    sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
    return qd, sz


def read_tl(x, ncta):
    ws = ops.debug_workspace(x)
    tl = ws[48 * 1024: 48 * 1024 + 148 * 64].view(torch.int64).reshape(148, 8)[:ncta].cpu()
    return tl


Ms = (1, 32) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(','))
for M in Ms:
    for (N, K) in [(14336, 4096)]:
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        wa, wb = mk(N, K), mk(N, K)
        for back_to_back in (False, True):
            for _ in range(2):
                ops.int4_tilepacked_linear(x, wa[0], g, wa[1], None, N, 1)
            torch.cuda.synchronize()
            ws = ops.debug_workspace(x)
            ws[48 * 1024: 48 * 1024 + 148 * 64].zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if back_to_back:
                ops.int4_tilepacked_linear(x, wa[0], g, wa[1], None, N, 1)
            ops.int4_tilepacked_linear(x, wb[0], g, wb[1], None, N, 1)
            e1.record()
            torch.cuda.synchronize()
            tl = read_tl(x, 148)
            used = (tl[:, 7] > 0)
            tl = tl[used]
            gt = tl[:, 0]
            ph = tl[:, 1:].float()
            med = ph.median(dim=0).values.tolist()
            mx = ph.max(dim=0).values.tolist()
            ws2 = ops.debug_workspace(x)
            fine = ws2[48 * 1024 + 148 * 64: 48 * 1024 + 148 * 64 + 4 * 64].view(torch.int64).reshape(4, 8).cpu().tolist()
            print(f"M={M:2d} N={N:5d} K={K:5d} b2b={int(back_to_back)} ctas={int(used.sum())} total={e0.elapsed_time(e1)*1e3:7.1f}us "
                  f"entry_spread={(gt.max()-gt.min()).item()/1e3:5.2f}us  med(cyc)={[int(v) for v in med]}  max={[int(v) for v in mx]}")
            if not back_to_back:
                print("     units 8..11 of CTA0 [wfull, math_done, aempty_ok, st_done, arrived, mma_start, mma_issued]:")
                for row in fine:
                    print("      ", row[:7])
