"""Absolute (globaltimer) phase timeline of the TS-mode GEMM over a chain of launches (AO_B200_TIMELINE=1).
Per CTA stamps: 0 entry, 1 prologue done, 2 pdl_wait returned, 3 first weights landed, 4 first MMA issued,
5 last MMA issued, 6 last accumulator ready, 7 epilogue done.  Two timeline slots alternate between launches, so the
last two kernels of a chain can be laid on one time axis: how much of kernel k+1 overlaps kernel k."""
import os
import sys

os.environ["AO_B200_TIMELINE"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.ops.load_library(os.path.join(ROOT, "ao_b200", "lib", "ao_b200_torch.so"))
ops = torch.ops.ao_b200
g = 32
SLOT = (100 * 16 + 64) * 8
BASE = 20 << 20


def mk(N, K):
    qd = torch.randint(-2**31, 2**31 - 1, (N // 8, K // 128, 32, 4), device="cuda", dtype=torch.int32)
    sz = (torch.rand(K // g, N, 2, device="cuda") * 0.01).to(torch.bfloat16)
    return qd, sz


def slot(ws, i):
    return ws[BASE + i * SLOT: BASE + i * SLOT + 100 * 16 * 8].view(torch.int64).reshape(100, 16).cpu()


def fine(ws, i):
    return ws[BASE + i * SLOT + 100 * 16 * 8: BASE + (i + 1) * SLOT].view(torch.int64).reshape(8, 8).cpu()


Ms = (1, 32) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(','))
shapes = [(14336, 4096)] if len(sys.argv) < 3 else [tuple(int(v) for v in sys.argv[2].split('x'))]
names = ["entry", "prologue", "pdl_wait", "w_landed", "mma_first", "mma_last", "acc_ready", "done", "flags_ok", "exit"]
for M in Ms:
    for (N, K) in shapes:
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        ws_list = [mk(N, K) for _ in range(4)]
        for _ in range(2):
            ops.int4_tilepacked_linear(x, ws_list[0][0], g, ws_list[0][1], None, N, 1)
        torch.cuda.synchronize()
        ws = ops.debug_workspace(x)
        ws[BASE: BASE + 2 * SLOT].zero_()
        torch.cuda.synchronize()
        for c in range(4):   # chain of 4; slots hold launches 2 and 3 (or 3 and 2)
            ops.int4_tilepacked_linear(x, ws_list[c][0], g, ws_list[c][1], None, N, 1)
        torch.cuda.synchronize()
        a, b = slot(ws, 0), slot(ws, 1)
        ua, ub = a[a[:, 9] > 0], b[b[:, 9] > 0]
        first, second = (ua, ub) if ua[:, 0].min() < ub[:, 0].min() else (ub, ua)
        t0 = int(first[:, 0].min())
        print(f"M={M} N={N} K={K}: chain of 4, last two kernels (us since the earlier one's first CTA entry; min / median / max over CTAs)")
        for nm, kern in (("k", first), ("k+1", second)):
            parts = []
            for e in range(10):
                col = (kern[:, e][kern[:, e] > 0] - t0).float() / 1e3
                if col.numel() == 0:
                    continue
                parts.append(f"{names[e]} {col.min():.1f}/{col.median():.1f}/{col.max():.1f}")
            print(f"   {nm:4s}" + "  ".join(parts))
        # the slowest CTAs of the later kernel and where they sit in the stream-K split
        KT, G = K // 128, 148
        Ntiles = (N + 127) // 128
        U = Ntiles * ((M + (16 if M <= 16 else 32 if M <= 32 else 64 if M <= 64 else 128) - 1) // (16 if M <= 16 else 32 if M <= 32 else 64 if M <= 64 else 128)) * KT
        sl = slot(ws, 0) if ua[:, 0].min() >= ub[:, 0].min() else slot(ws, 1)
        order = sorted(range(100), key=lambda i: -int(sl[i, 9]))[:6]
        med = sorted(int(sl[i, 9]) for i in range(100) if int(sl[i, 9]) > 0)[50]
        for bidx in order:
            u0, u1 = U * bidx // G, U * (bidx + 1) // G
            kc0, n = u0 % KT, u1 - u0
            first = min(KT - kc0, n)
            segs = [("contrib" if kc0 else ("full" if first == KT else "owner"), first)]
            rest = n - first
            while rest > 0:
                c = min(KT, rest)
                segs.append(("full" if c == KT else "owner", c))
                rest -= c
            row = "  ".join(f"{names[e]} {(int(sl[bidx, e]) - t0) / 1e3:.1f}" for e in (4, 5, 6, 8, 9) if int(sl[bidx, e]) > 0)
            print(f"   slow CTA {bidx:3d} (+{(int(sl[bidx, 9]) - med) / 1e3:.1f} us vs median exit): units {n} segments {segs}: {row}")
        f = fine(ws, 0) if ua[:, 0].min() >= ub[:, 0].min() else fine(ws, 1)   # the later kernel's CTA 0
        if int(f[0, 0]) > 0:
            base = int(f[0, 0])
            print("   CTA 0, chunks 16..23 (SM cycles since chunk 16 was requested): requested / landed / stage back / A stored / issuer saw / committed / A free")
            for c in range(8):
                print("     chunk %2d: " % (16 + c) + "  ".join("%6d" % (int(f[c, e]) - base) if int(f[c, e]) > 0 else "     -" for e in range(7)))
        print(f"   launch-to-launch: {(int(second[:, 9].max()) - int(first[:, 9].max())) / 1e3:.2f} us  (end of k+1 minus end of k)")
