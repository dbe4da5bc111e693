"""Index mathematics of the tcgen05 convolution kernels (csrc/conv_tc.cu), restated in numpy and checked
on the CPU: the GEMM formulations (implicit-GEMM forward, parity-class input gradient, taps x positions
weight gradient) against torch's convolution derivatives in fp64, and the producer thread -> shared-memory
mappings (SWIZZLE_128B chunk placement) as exact covers of the operand tiles.  These are the host-side
invariants the CUDA code relies on; the kernels themselves are tested on the GPU (test_gpu_gemm.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def im2col(x, k, s, p):
    """[N,C,H,W] -> A[m = (n,oy,ox), tap = (c,ky,kx)] - the GEMM row/column order of the kernels."""
    N, C, H, W = x.shape
    cols = F.unfold(x, kernel_size=k, stride=s, padding=p)          # [N, C*k*k, P], taps ordered (c,ky,kx)
    return cols.permute(0, 2, 1).reshape(-1, C * k * k)


@pytest.mark.parametrize("geom", [dict(C=4, OC=16, k=8, s=4, p=0, H=36, W=44), dict(C=16, OC=32, k=4, s=2, p=1, H=20, W=20),
                                  dict(C=16, OC=32, k=4, s=2, p=1, H=7, W=5)])
def test_forward_and_wgrad_gemm_formulations(geom):
    g = torch.Generator().manual_seed(0)
    C, OC, k, s, p, H, W = (geom[n] for n in "C OC k s p H W".split())
    x = torch.randn(3, C, H, W, dtype=torch.float64, generator=g)
    w = torch.randn(OC, C, k, k, dtype=torch.float64, generator=g, requires_grad=True)
    y = F.conv2d(x, w, stride=s, padding=p)
    A = im2col(x, k, s, p)                                            # [M, K]
    Y = A @ w.detach().reshape(OC, -1).T                              # forward: rows = positions
    np.testing.assert_allclose(Y.reshape(3, -1, OC).permute(0, 2, 1).reshape(y.shape).numpy(), y.detach().numpy(),
                               rtol=1e-12, atol=1e-12)
    go = torch.randn(y.shape, dtype=torch.float64, generator=g)
    y.backward(go)
    G = go.permute(0, 2, 3, 1).reshape(-1, OC)                        # g[m, oc]
    D = A.T @ G                                                       # wgrad: D[tap, oc], reduction over positions
    np.testing.assert_allclose(D.T.reshape(w.shape).numpy(), w.grad.numpy(), rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("H,W", [(20, 20), (25, 19), (7, 5), (2, 2)])
def test_dgrad_parity_decomposition(H, W):
    """Input gradient of Conv2d(16->32,k4,s2,p1) as four GEMMs, one per input-pixel parity (py,px): pixel
    (2a+py, 2b+px) receives taps ky in {1,3} (py=0) or {0,2} (py=1) from output rows a+dpos - the
    Dgrad2 policy's ``dpos`` table - and likewise in x; K = 32 channels x 2 x 2 taps."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 16, H, W, dtype=torch.float64, generator=g, requires_grad=True)
    w = torch.randn(32, 16, 4, 4, dtype=torch.float64, generator=g)
    y = F.conv2d(x, w, stride=2, padding=1)
    go = torch.randn(y.shape, dtype=torch.float64, generator=g)
    y.backward(go)
    OH, OW = y.shape[2:]
    dpos = lambda parity, t: (0 if t == 0 else -1) if parity == 0 else (1 if t == 0 else 0)
    ktap = lambda parity, t: (1 if t == 0 else 3) if parity == 0 else (0 if t == 0 else 2)
    dx = torch.zeros_like(x)
    gon, wn = go.numpy(), w.numpy()
    out = dx.numpy()
    for py in range(2):
        for px in range(2):
            for iy in range(py, H, 2):
                for ix in range(px, W, 2):
                    a, b = iy // 2, ix // 2
                    acc = np.zeros((2, 16))
                    for ty in range(2):
                        for tx in range(2):
                            oy, ox = a + dpos(py, ty), b + dpos(px, tx)
                            if 0 <= oy < OH and 0 <= ox < OW:
                                acc += gon[:, :, oy, ox] @ wn[:, :, ktap(py, ty), ktap(px, tx)]
                    out[:, :, iy, ix] = acc
    np.testing.assert_allclose(out, x.grad.numpy(), rtol=1e-12, atol=1e-12)


def swizzled(row, chunk):
    """Byte offset of 16-byte chunk ``chunk`` of row ``row`` in a K-major SWIZZLE_128B tile (128 B rows)."""
    return row * 128 + ((chunk ^ (row & 7)) << 4)


def test_forward_producer_mapping_covers_tile():
    """conv_fwd_tc_kernel: thread -> (row r = tid & 127, chunks jh..jh+3 with jh = (tid >> 7) * 4)."""
    seen = set()
    for tid in range(256):
        r, jh = tid & 127, (tid >> 7) * 4
        for q in range(4):
            seen.add(swizzled(r, jh + q))
    assert seen == {16 * i for i in range(128 * 8)}


@pytest.mark.parametrize("kN", [16, 32])
def test_wgrad_producer_mapping_covers_tiles(kN):
    """conv_wgrad_tc_kernel: thread -> (4-position chunk mc, tap group tg); A^T rows are taps, G^T rows are
    output channels; every 16-byte (A^T) / kGP*4-byte (G^T) slot is written exactly once, and the eight
    lanes of one tap group fill one whole 128-byte row (conflict-free stores)."""
    kGP = kN // 8
    a_slots, g_slots = [], []
    for warp in range(8):
        for lane in range(32):
            mc = (lane >> 1) & 7
            tg = (lane & 1) + 2 * (lane >> 4) + 4 * warp
            for i in range(4):
                a_slots.append(swizzled(4 * tg + i, mc))
            g_oc, g_r0 = (tg * kGP) >> 2, (tg * kGP) & 3
            base = swizzled(g_oc, mc) + 4 * g_r0
            g_slots += [base + 4 * j for j in range(kGP)]
            assert 0 <= g_oc < kN and g_r0 + kGP <= 4
    assert sorted(a_slots) == [16 * i for i in range(128 * 8)]                 # one tap half: 128 rows x 8 chunks
    assert sorted(g_slots) == [4 * i for i in range(kN * 32)]                  # kN rows x 32 positions
    # lanes of a quarter-warp store phase (8 lanes) hit 8 distinct 16-byte bank groups
    for warp in range(8):
        for phase in range(4):
            lanes = range(8 * phase, 8 * phase + 8)
            offs = {swizzled(4 * ((l & 1) + 2 * (l >> 4) + 4 * warp), (l >> 1) & 7) % 128 for l in lanes}
            assert len(offs) == 8


def test_wgrad_position_walk_matches_division():
    """The carry-only (image, oy, ox) walk of the wgrad producers equals m -> (m // P, (m % P) // OW, m % OW)
    for every k-block, chunk and row, including P < 32 and P = 1."""
    for OH, OW in [(20, 20), (25, 19), (10, 10), (2, 3), (1, 1)]:
        P = OH * OW
        adv_y, adv_x = 32 // OW, 32 - (32 // OW) * OW
        for mc in range(8):
            for kb_begin in (0, 5):
                m0 = kb_begin * 32 + 4 * mc
                n, pos = divmod(m0, P)
                oy, ox = divmod(pos, OW)
                for kb in range(kb_begin, kb_begin + 40):
                    nn, yy, xx = n, oy, ox
                    for r in range(4):
                        m = kb * 32 + 4 * mc + r
                        assert (nn, yy, xx) == (m // P, (m % P) // OW, m % OW)
                        xx += 1
                        if xx == OW:
                            xx = 0
                            yy += 1
                            if yy == OH:
                                yy = 0
                                nn += 1
                    ox += adv_x
                    oy += adv_y
                    if ox >= OW:
                        ox -= OW
                        oy += 1
                    while oy >= OH:
                        oy -= OH
                        n += 1


# ------------------------------------------------------------------------------------------------------
# Planned v2 formulation: space-to-depth turns both strided convolutions into
# 2x2 stride-1 convolutions over 64 channels, so the four taps are four ROW-SHIFTED views of one operand
# tile (shifted tcgen05 descriptors) and every input element is converted / written to shared memory
# once instead of four times.  These tests pin the identities the kernels will rely on.
def space_to_depth(x, s, pad):
    """[N,C,H,W] -> X[N, Y*GW + X, (c, ky', kx')] with Y = (H + 2 pad)/s rows of the padded grid."""
    N, C, H, W = x.shape
    xp = F.pad(x, (pad, pad, pad, pad))
    GH, GW = (H + 2 * pad) // s, (W + 2 * pad) // s
    xp = xp[:, :, :GH * s, :GW * s].reshape(N, C, GH, s, GW, s)
    return xp.permute(0, 2, 4, 1, 3, 5).reshape(N, GH * GW, C * s * s), GH, GW


@pytest.mark.parametrize("geom", [dict(C=4, OC=16, k=8, s=4, p=0, H=84, W=84), dict(C=16, OC=32, k=4, s=2, p=1, H=20, W=20),
                                  dict(C=4, OC=16, k=8, s=4, p=0, H=104, W=80)])
def test_space_to_depth_shifted_gemm(geom):
    g = torch.Generator().manual_seed(2)
    C, OC, k, s, p, H, W = (geom[n] for n in "C OC k s p H W".split())
    N = 2
    x = torch.randn(N, C, H, W, dtype=torch.float64, generator=g)
    w = torch.randn(OC, C, k, k, dtype=torch.float64, generator=g, requires_grad=True)
    y = F.conv2d(x, w, stride=s, padding=p)
    OH, OW = y.shape[2:]
    X, GH, GW = space_to_depth(x, s, p)                               # [N, GH*GW, 64]
    assert X.shape[2] == 64 and GH >= OH + 1 and GW >= OW + 1
    # W4[(by,bx)][oc, (c,ky',kx')] = w[oc, c, s*by+ky', s*bx+kx']
    W4 = w.detach().reshape(OC, C, 2, s, 2, s).permute(2, 4, 0, 1, 3, 5).reshape(2, 2, OC, C * s * s)
    flat = X.reshape(N * GH * GW, 64)                                 # images back to back, as in HBM
    rows = torch.arange(N * GH * GW)
    Y = torch.zeros(N * GH * GW, OC, dtype=torch.float64)
    for by in range(2):
        for bx in range(2):
            shift = by * GW + bx                                      # the descriptor row offset of this tap
            src = torch.clamp(rows + shift, max=N * GH * GW - 1)      # rows past the end feed discarded outputs only
            Y += flat[src] @ W4[by, bx].T
    Yg = Y.reshape(N, GH, GW, OC)[:, :OH, :OW].permute(0, 3, 1, 2)     # valid grid positions only
    np.testing.assert_allclose(Yg.numpy(), y.detach().numpy(), rtol=1e-11, atol=1e-11)
    # weight gradient with the SMALL operand shifted instead: D[(by,bx)][ch, oc] = sum_pos X[pos, ch] * G[pos - shift, oc]
    go = torch.randn(y.shape, dtype=torch.float64, generator=g)
    y.backward(go)
    Gg = torch.zeros(N, GH, GW, OC, dtype=torch.float64)
    Gg[:, :OH, :OW] = go.permute(0, 2, 3, 1)                           # zero on the invalid last row / column
    Gf = Gg.reshape(N * GH * GW, OC)
    dW4 = torch.zeros(2, 2, OC, C * s * s, dtype=torch.float64)
    for by in range(2):
        for bx in range(2):
            shift = by * GW + bx
            Gs = torch.zeros_like(Gf)
            Gs[shift:] = Gf[:Gf.shape[0] - shift] if shift else Gf
            dW4[by, bx] = (flat.T @ Gs).T
    dW = dW4.reshape(2, 2, OC, C, s, s).permute(2, 3, 0, 4, 1, 5).reshape(w.shape)
    np.testing.assert_allclose(dW.numpy(), w.grad.numpy(), rtol=1e-10, atol=1e-10)


def test_v2_probe_index_mappings():
    """Thread-level emulation of tools/probes/conv1_v2_probe.cu (items -> stage rows/chunks, filter-bank
    tiles, shifted tap reads, epilogue decode) against torch's convolution - everything of the probe
    except the hardware descriptor semantics, which only the GPU run can answer."""
    N, H, W = 3, 84, 84
    GH, GW, OH, OW = H // 4, W // 4, 20, 20
    G, g_total = GH * GW, N * GH * GW
    kRows, kStageRows, kProd = 128, 160, 256
    gen = torch.Generator().manual_seed(5)
    x = torch.randint(0, 256, (N, 4, H, W), dtype=torch.uint8, generator=gen)
    w = torch.randn(16, 4, 8, 8, dtype=torch.float64, generator=gen) / 8
    b = torch.randn(16, dtype=torch.float64, generator=gen)
    xf = x.reshape(-1).numpy()
    # filter bank tiles: Bt[tap][kb][oc][32]
    Bt = np.zeros((4, 2, 16, 32))
    for idx in range(4 * 2 * 16 * 8):
        j, oc, kb, tap = idx & 7, (idx >> 3) & 15, (idx >> 7) & 1, idx >> 8
        by, bx = tap >> 1, tap & 1
        ch0 = kb * 32 + j * 4
        c, kyp = ch0 >> 4, (ch0 >> 2) & 3
        Bt[tap, kb, oc, 4 * j:4 * j + 4] = w[oc, c, 4 * by + kyp, 4 * bx:4 * bx + 4].numpy()
    out = np.full((N, 16, OH, OW), np.nan)
    for tile in range((g_total + kRows - 1) // kRows):
        # row table (one decode per stage row)
        row_off = np.full(kStageRows, -1, dtype=np.int64)
        for p in range(kStageRows):
            gg = tile * kRows + p
            if gg < g_total:
                n, pos = divmod(gg, G)
                Yg, Xg = divmod(pos, GW)
                row_off[p] = n * 4 * H * W + 4 * Yg * W + 4 * Xg
        stage = np.zeros((2, kStageRows, 32))
        for tid in range(kProd):
            for i in range(kStageRows * 16 // kProd):
                item = tid + i * kProd
                p, q = item % kStageRows, item // kStageRows
                c, kyp = q >> 2, q & 3
                if row_off[p] >= 0:
                    a = row_off[p] + (c * H + kyp) * W
                    stage[q >> 3, p, 4 * (q & 7):4 * (q & 7) + 4] = xf[a:a + 4]
        acc = np.zeros((2, kRows, 16))
        for tap in range(4):
            shift = (tap >> 1) * GW + (tap & 1)
            for kb in range(2):
                acc[tap >> 1] += stage[kb, shift:shift + kRows] @ Bt[tap, kb].T
        for i in range(kRows):
            gg = tile * kRows + i
            if gg < g_total:
                n, pos = divmod(gg, G)
                Yg, Xg = divmod(pos, GW)
                if Yg < OH and Xg < OW:
                    out[n, :, Yg, Xg] = np.maximum((acc[0, i] + acc[1, i]) / 255.0 + b.numpy(), 0)
    ref = F.relu(F.conv2d(x.double() / 255, w, b, stride=4)).numpy()
    assert not np.isnan(out).any()
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-10)


def test_wgrad_v2_probe_index_mappings():
    """Thread-level emulation of tools/probes/conv1_wgrad_v2_probe.cu: K-block = grid row, A^T units (four
    words -> four transposed chunks), G units written aligned (bx = 0) and one element to the right (bx = 1),
    the [64 ch x 64 (tap, oc)] accumulator and the reduce kernel's scatter into the [16,4,8,8] gradient."""
    N, H, W = 2, 84, 84
    GH, GW, OH, OW = H // 4, W // 4, 20, 20
    n16, nG, P = (GW + 3) // 4, (OW + 3) // 4, OH * OW
    gen = torch.Generator().manual_seed(6)
    x = torch.randint(0, 256, (N, 4, H, W), dtype=torch.uint8, generator=gen)
    w = torch.zeros(16, 4, 8, 8, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(16, dtype=torch.float64, requires_grad=True)
    out = torch.randn(N, 16, OH, OW, dtype=torch.float64, generator=gen)
    go = torch.randn(N, 16, OH, OW, dtype=torch.float64, generator=gen)
    gm = go * (out > 0)
    F.conv2d(x.double() / 255, w, b, stride=4).backward(gm)
    xf, gf, of = x.reshape(-1).numpy(), go.reshape(-1).numpy(), out.reshape(-1).numpy()
    D = np.zeros((64, 64))
    bias = np.zeros(16)
    for kb in range(N * GH):
        n, Yg = divmod(kb, GH)
        At = np.zeros((64, 32))
        Bt = np.zeros((64, 32))
        for tid in range(16 * n16):                                   # A^T units
            a_c, a_ky, a_j = tid // (4 * n16), (tid // n16) & 3, tid % n16
            words = min(4, (W - a_j * 16) // 4)
            base = n * 4 * H * W + (a_c * H + 4 * Yg + a_ky) * W + a_j * 16
            for pos in range(words):
                for kx in range(4):
                    At[a_c * 16 + a_ky * 4 + kx, 4 * a_j + pos] = xf[base + 4 * pos + kx]
        for gt in range(32 * nG):                                     # G units
            g_row, g_j = divmod(gt, nG)
            g_oc, g_by = g_row & 15, (g_row >> 4) & 1
            oy = Yg - g_by
            e = np.zeros(4)
            if 0 <= oy < OH:
                base = (n * 16 + g_oc) * P + oy * OW + 4 * g_j
                for t in range(4):
                    if 4 * g_j + t < OW:
                        e[t] = gf[base + t] if of[base + t] > 0 else 0.0
            if g_by == 0:
                bias[g_oc] += e.sum()
            row0 = (g_by * 2) * 16 + g_oc
            Bt[row0, 4 * g_j:4 * g_j + 4] = e
            for t in range(4):
                Bt[row0 + 16, 4 * g_j + t + 1] = e[t]
        D += At @ Bt.T
    dW = np.zeros((16, 4, 8, 8))
    for i in range(64 * 64):                                          # wgrad_v2_reduce_kernel
        ch, col = i >> 6, i & 63
        tap, oc = col >> 4, col & 15
        by, bx = tap >> 1, tap & 1
        c, kyp, kxp = ch >> 4, (ch >> 2) & 3, ch & 3
        dW[oc, c, 4 * by + kyp, 4 * bx + kxp] = D[ch, col] / 255.0
    np.testing.assert_allclose(dW, w.grad.numpy(), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(bias, b.grad.numpy(), rtol=1e-10, atol=1e-10)
