"""GPU: per-kernel unit tests of the sm_100a kernels against plain torch fp32 ops / the oracle."""
import pytest
import torch
import torch.nn.functional as F

from popcorn_b200 import _lib, ops
from oracle import popcorn_oracle as po

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _st():
    return torch.cuda.current_stream().cuda_stream


def _pack_conv(w, b):
    return torch.cat([w.permute(1, 2, 3, 0).reshape(-1), b]).contiguous()


def _pack_conv_tc(w, b, tc):
    """tc=True: the layer's tcgen05 weight image (hi/lo split operands, SWIZZLE_128B), built by the library's host packer."""
    if not tc:
        return None
    L = _lib.lib()
    cout, cin = w.shape[:2]
    flat = _pack_conv(w, b).cpu()
    img = torch.zeros(L.pc_conv_tc_layer_floats(cin, cout), dtype=torch.float32)
    _lib.check(L.pc_conv_tc_pack_layer(flat.data_ptr(), cin, cout, img.data_ptr()))
    return img.cuda()


def _p(t):
    return None if t is None else t.data_ptr()


TC = pytest.mark.parametrize("tc", [False, True], ids=["simt", "tcgen05"])


@pytest.mark.parametrize("cin,cout", [(2, 8), (4, 8), (8, 8), (8, 16), (16, 16)])
@pytest.mark.parametrize("H,W", [(64, 64), (37, 53), (32, 100), (70, 300)])
@TC
def test_conv3x3_bias_relu(cin, cout, H, W, tc):
    g = torch.Generator().manual_seed(cin * 100 + cout + H)
    x = torch.randn(cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.3).cuda()
    b = torch.randn(cout, generator=g).cuda()
    out = torch.full((cout, H, W), float("nan"), device="cuda")
    wtc = _pack_conv_tc(w, b, tc)
    _lib.check(_lib.lib().pc_test_conv3x3(x.data_ptr(), cin, H, W, 0, 0, 0, None, 0, 0, 0, 0, 0,
                                          _pack_conv(w, b).data_ptr(), cout, H, W, out.data_ptr(), None, _p(wtc), _st()))
    ref = F.relu(F.conv2d(x[None].cpu(), w.cpu(), b.cpu(), padding=1))[0]
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-4), float((out.cpu() - ref).abs().max())


@pytest.mark.parametrize("scale,rel", [(1e-4, 1e-3), (1.0, 5e-6), (2e4, 5e-6), (1e5, 3e-4)])
def test_conv3x3_tc_operand_range(scale, rel):
    """Split-operand range of the tensor-core convs (csrc/tc_common.cuh split_f16x2): fp32-like accuracy for O(1) .. O(1e4) activations,
    graceful for tiny ones (fp16 halves: lo parts go subnormal, absolute error ~3e-8 per element) and inside the saturating window
    65504 < |x| <= 131008; the TF32 build (pc_tc_operand_format() == 0) has no such window and meets the tight bar everywhere."""
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(8, 40, 160, generator=g) * scale).cuda()           # non-negative like post-ReLU activations, max = scale
    w = (torch.randn(8, 8, 3, 3, generator=g) * 0.3).cuda()
    b = torch.zeros(8).cuda()
    out = torch.full((8, 40, 160), float("nan"), device="cuda")
    wtc = _pack_conv_tc(w, b, True)
    _lib.check(_lib.lib().pc_test_conv3x3(x.data_ptr(), 8, 40, 160, 0, 0, 0, None, 0, 0, 0, 0, 0, _pack_conv(w, b).data_ptr(), 8, 40, 160,
                                          out.data_ptr(), None, _p(wtc), _st()))
    ref = F.conv2d(x[None].double().cpu(), w.double().cpu(), None, padding=1)[0].clamp_(min=0)
    err = float((out.double().cpu() - ref).abs().max() / ref.abs().max())
    bar = rel if _lib.lib().pc_tc_operand_format() == 1 else 5e-6
    assert torch.isfinite(out).all() and err < bar, err


def test_conv3x3_tc_overflow_is_loud():
    """fp16 halves: an activation beyond 2 x 65504 cannot be represented by hi + lo; the lo half overflows to inf on purpose, so the pixels
    that see it come out non-finite instead of silently clipped (fp32 / the TF32 build compute them)."""
    x = torch.ones(8, 32, 128, device="cuda")
    x[3, 10, 40] = 3.0e5
    w = torch.full((8, 8, 3, 3), 0.1, device="cuda")
    b = torch.zeros(8, device="cuda")
    out = torch.zeros(8, 32, 128, device="cuda")
    wtc = _pack_conv_tc(w, b, True)
    _lib.check(_lib.lib().pc_test_conv3x3(x.data_ptr(), 8, 32, 128, 0, 0, 0, None, 0, 0, 0, 0, 0, _pack_conv(w, b).data_ptr(), 8, 32, 128,
                                          out.data_ptr(), None, _p(wtc), _st()))
    hit = out[:, 9:12, 39:42]
    if _lib.lib().pc_tc_operand_format() == 1:
        assert not torch.isfinite(hit).any()
        keep = torch.ones_like(out, dtype=torch.bool)
        keep[:, 9:12, 39:42] = False
        assert torch.isfinite(out[keep]).all()
    else:
        assert torch.isfinite(out).all() and abs(float(hit[0, 1, 1]) - (7.1 + 3.0e4)) < 0.1


@pytest.mark.parametrize("stream", ["sar", "optical"])
@pytest.mark.parametrize("H,W,tc_path", [(64, 128, True), (70, 300, True), (37, 53, False)], ids=["64x128", "70x300", "odd_falls_back"])
def test_first_layer_on_tensor_cores(stream, H, W, tc_path):
    """The first conv layer of a stream on tcgen05 (csrc/conv.cu launch_conv, first_layer): an unpadded source whose channel map is a
    contiguous run of planes — SAR = planes (4, 5) of a 6-plane tensor, optical = planes (2, 1, 0, 3) with the permutation folded into the
    weight image — is read with ONE TMA box of Cin planes; anything else (odd strides here) falls back to the fp32 stencil.  Both paths
    must agree with torch's conv on the reordered channels."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(6, H, W, generator=g).cuda()
    cin = 2 if stream == "sar" else 4
    order = [4, 5] if stream == "sar" else [2, 1, 0, 3]
    chmap = 0x00000504 if stream == "sar" else 0x03000102
    w = (torch.randn(8, cin, 3, 3, generator=g) * 0.3)
    b = torch.randn(8, generator=g)
    flat = _pack_conv(w, b)                                   # network channel order (what the fp32 stencil multiplies)
    w_mem = w if stream == "sar" else w[:, [2, 1, 0, 3]]      # memory plane order (what the tensor-core image holds): plane q = logical (2,1,0,3)[q]
    img = torch.zeros(L.pc_conv_tc_layer_floats(cin, 8))
    flat_mem = _pack_conv(w_mem, b)                           # named: the host packer reads it through a raw pointer
    _lib.check(L.pc_conv_tc_pack_layer(flat_mem.data_ptr(), cin, 8, img.data_ptr()))
    img, flat_d = img.cuda(), flat.cuda()
    out = torch.full((8, H, W), float("nan"), device="cuda")
    ops.profile_enable(True)
    _lib.check(L.pc_conv3x3_layer(x.data_ptr(), cin, x.stride(0), x.stride(1), H, W, 0, 0, 0, chmap, None, 0, 0, 0, 0, 0, 0, 0,
                                  flat_d.data_ptr(), img.data_ptr(), 8, 1, H, W, out.data_ptr(), out.stride(0), out.stride(1),
                                  None, 0, 0, _st()), "pc_conv3x3_layer")
    torch.cuda.synchronize()
    ops.profile_enable(False)
    ran = [k for k, v in ops.profile_results().items() if v[1] > 0]
    assert any(k.startswith(f"conv3x3_tc<{cin},0,8") for k in ran) == tc_path, ran
    ref = F.relu(F.conv2d(x[order][None].cpu(), w, b, padding=1))[0]
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-4), float((out.cpu() - ref).abs().max())


@pytest.mark.parametrize("c", [8, 16])
@pytest.mark.parametrize("H,W", [(64, 64), (37, 52), (50, 36), (66, 260)])
@TC
def test_conv3x3_fused_maxpool(c, H, W, tc):
    g = torch.Generator().manual_seed(c + H)
    x = torch.randn(c, H, W, generator=g).cuda()
    w = (torch.randn(c, c, 3, 3, generator=g) * 0.3).cuda()
    b = torch.randn(c, generator=g).cuda()
    out = torch.empty(c, H, W, device="cuda")
    pool = torch.full((c, H // 2, W // 2), float("nan"), device="cuda")
    wtc = _pack_conv_tc(w, b, tc)
    _lib.check(_lib.lib().pc_test_conv3x3(x.data_ptr(), c, H, W, 0, 0, 0, None, 0, 0, 0, 0, 0,
                                          _pack_conv(w, b).data_ptr(), c, H, W, out.data_ptr(), pool.data_ptr(), _p(wtc),
                                          _st()))
    ref = F.relu(F.conv2d(x[None].cpu(), w.cpu(), b.cpu(), padding=1))
    assert torch.allclose(out.cpu(), ref[0], rtol=1e-4, atol=1e-4)
    assert torch.allclose(pool.cpu(), F.max_pool2d(ref, 2)[0], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("ca,cb,cout", [(16, 16, 8), (8, 8, 8)])
@pytest.mark.parametrize("H,W", [(64, 64), (37, 53), (41, 259)])
@TC
def test_conv3x3_concat_with_offset_zero_pad(ca, cb, cout, H, W, tc):
    """Up block: cat[skip, F.pad(upsampled)] -> conv (networks.py:309-319) without materialising either."""
    g = torch.Generator().manual_seed(ca + H)
    a = torch.randn(ca, H, W, generator=g).cuda()
    bH, bW = 2 * (H // 2), 2 * (W // 2)
    bsrc = torch.randn(cb, bH, bW, generator=g).cuda()
    w = (torch.randn(cout, ca + cb, 3, 3, generator=g) * 0.2).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    dy, dx = H - bH, W - bW
    out = torch.empty(cout, H, W, device="cuda")
    wtc = _pack_conv_tc(w, bias, tc)
    _lib.check(_lib.lib().pc_test_conv3x3(a.data_ptr(), ca, H, W, 0, 0, 0, bsrc.data_ptr(), cb, bH, bW, dy // 2, dx // 2,
                                          _pack_conv(w, bias).data_ptr(), cout, H, W, out.data_ptr(), None, _p(wtc), _st()))
    bp = F.pad(bsrc.cpu(), (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    ref = F.relu(F.conv2d(torch.cat([a.cpu(), bp])[None], w.cpu(), bias.cpu(), padding=1))[0]
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-4), float((out.cpu() - ref).abs().max())


@pytest.mark.parametrize("pad", [(14, 14), (5, 0), (0, 9)])
@TC
def test_conv3x3_virtual_reflect_padding(pad, tc):
    """Reflect padding folded into the loader (popcorn.py:244, 292): conv over the virtually padded image."""
    g = torch.Generator().manual_seed(3)
    H, W, cin, cout = 45, 161, 4, 8
    py, px = pad
    x = torch.randn(cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.3).cuda()
    b = torch.randn(cout, generator=g).cuda()
    Hv, Wv = H + 2 * py, W + 2 * px
    out = torch.empty(cout, Hv, Wv, device="cuda")
    wtc = _pack_conv_tc(w, b, tc)
    _lib.check(_lib.lib().pc_test_conv3x3(x.data_ptr(), cin, H, W, py, px, 1, None, 0, 0, 0, 0, 0,
                                          _pack_conv(w, b).data_ptr(), cout, Hv, Wv, out.data_ptr(), None, _p(wtc), _st()))
    xp = F.pad(x[None].cpu(), (px, px, py, py), mode="reflect")
    ref = F.relu(F.conv2d(xp, w.cpu(), b.cpu(), padding=1))[0]
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("c", [8, 16])
def test_conv_transpose_2x2(c):
    g = torch.Generator().manual_seed(c)
    Hl, Wl = 19, 45
    x = torch.randn(c, Hl, Wl, generator=g).cuda()
    w = (torch.randn(c, c, 2, 2, generator=g) * 0.3).cuda()
    b = torch.randn(c, generator=g).cuda()
    pack = torch.cat([w.permute(0, 2, 3, 1).reshape(-1), b]).contiguous()
    out = torch.empty(c, 2 * Hl, 2 * Wl, device="cuda")
    _lib.check(_lib.lib().pc_test_convt2x2(x.data_ptr(), c, Hl, Wl, pack.data_ptr(), out.data_ptr(), _st()))
    ref = F.conv_transpose2d(x[None].cpu(), w.cpu(), b.cpu(), stride=2)[0]
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("npix,R", [(1000, 7), (128 * 1024 + 37, 400), (3_000_001, 72976), (0, 3)])
def test_region_sum_and_backward(npix, R):
    g = torch.Generator().manual_seed(npix % 1000 + R)
    dens = torch.rand(npix, generator=g)
    # spatially coherent runs plus a noisy stretch, background 0 and out-of-range ids (-1, R) that must be ignored
    run = torch.randint(1, 5000, (max(1, npix // 700 + 2),), generator=g)
    ids = torch.repeat_interleave(torch.randint(-1, R + 1, (len(run),), generator=g), run)[:npix].to(torch.int32)
    if npix > 5000:
        ids[1000:3000] = torch.randint(0, R, (2000,), generator=g).to(torch.int32)
    sums = ops.region_sum(dens.cuda(), ids.cuda(), R).cpu()
    valid = (ids >= 0) & (ids < R)
    ref = torch.zeros(R, dtype=torch.float64).index_add_(0, ids[valid].long(), dens[valid].double())
    assert torch.allclose(sums, ref, rtol=1e-5, atol=1e-6)
    if npix:
        gs = torch.randn(R, generator=g)
        gd = ops.region_sum_backward(gs.cuda(), ids.cuda()).cpu()
        refg = torch.where(valid, gs[ids.clamp(0, R - 1).long()], torch.zeros(()))
        assert torch.equal(gd, refg)
        fac = torch.rand(R, generator=g) + 0.5
        d2 = ops.region_scale_(dens.clone().cuda(), ids.cuda(), fac.cuda()).cpu()
        assert torch.equal(d2, torch.where(valid, dens * fac[ids.clamp(0, R - 1).long()], dens))


@pytest.mark.parametrize("B,H,W", [(2, 72, 88), (1, 300, 517), (3, 33, 41)])
@pytest.mark.parametrize("empty", [False, True])
def test_sparsity_mask_compaction_is_bit_exact(B, H, W, empty):
    g = torch.Generator().manual_seed(B * H + W)
    builtup = torch.rand(B, 1, H, W, generator=g)
    builtup[:, :, ::3] = 0.0                                    # exercise the (builtup > 0) term
    admin = torch.randint(0, 4, (B, H, W), generator=g).float()
    admin[:, :, -5:] = -1.0
    cidx = torch.tensor([2, 3, 1][:B])
    if empty:
        admin[admin == cidx.view(-1, 1, 1).float()] = 0.0       # no pixel of the region -> fallback branch (:374-375)
    torch.manual_seed(11)
    grid = po.sparsity_grid(H, W)
    ref = po.sparsity_mask(builtup, admin, cidx, True, grid)
    rows = torch.zeros(H, dtype=torch.uint8); rows[grid[0]] = 1
    cols = torch.zeros(W, dtype=torch.uint8); cols[grid[1]] = 1
    mask, idx, n = ops.sparse_mask_compact(builtup[:, 0].cuda(), admin.cuda(), cidx.int().cuda(), rows.cuda(), cols.cuda())
    n = int(n.item())
    assert torch.equal(mask.bool().cpu(), ref)                                   # index set bit-exact
    assert torch.equal(idx[:n].long().cpu(), ref.reshape(-1).nonzero()[:, 0])    # row-major (b,h,w) order


@pytest.mark.parametrize("H,W,ps,ov", [(200, 236, 96, 16), (201, 237, 97, 15)], ids=["aligned_vector_path", "odd_scalar_path"])
def test_accumulate_and_finalize_match_run_eval_arithmetic(H, W, ps, ov):
    g = torch.Generator().manual_seed(1)
    maps = [torch.zeros(H, W, device="cuda") for _ in range(4)] + [torch.zeros(H, W, dtype=torch.int16, device="cuda")]
    ref = [torch.zeros(H, W) for _ in range(4)] + [torch.zeros(H, W, dtype=torch.int16)]
    m = po.centre_mask(ps, ps, ov)
    for xl, yl in po.get_patch_indices(H, W, ps, ov).tolist():
        d = torch.rand(ps, ps, generator=g)
        s = torch.rand(ps, ps, generator=g)
        ops.accumulate_tile(d.cuda(), s.cuda(), (ov, ps - ov), (ov, ps - ov), maps, xl, yl)
        for t, v in zip(ref[:4], (d, d * d, s, s * s)):
            t[xl:xl + ps, yl:yl + ps][m] += v[m]
        ref[4][xl:xl + ps, yl:yl + ps][m] += 1
    ops.finalize_map(maps)
    div = ref[4] > 1
    cf = ref[4][div].float()
    ref[0][div] = ref[0][div] / cf
    ref[1][div] = torch.sqrt((ref[1][div] - ref[0][div] ** 2 * cf) / (cf - 1))
    ref[2][div] = ref[2][div] / cf
    ref[3][div] = torch.sqrt((ref[3][div] - ref[2][div] ** 2 * cf) / (cf - 1))
    assert torch.equal(maps[4].cpu(), ref[4]) and int(div.sum()) > 0
    for a, b in zip(maps[:4], ref[:4]):
        assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=1e-5, equal_nan=True)


@pytest.mark.parametrize("cols,origin", [((16, 78), (8, 12)), ((16, 80), (8, 12)), ((17, 80), (8, 12)), ((16, 80), (8, 13))],
                         ids=["vector_with_tail", "vector", "odd_c0", "odd_x0"])
def test_accumulate_vector_and_scalar_paths_agree_with_torch(cols, origin):
    """pc_accumulate_tile takes 4 pixels per thread when tile, maps and column range are 16-byte aligned (with a scalar tail) and one pixel
    per thread otherwise: every variant must give the element-wise result bit for bit; pc_finalize_map reads the counts 8 at a time when
    its slice is 16-byte aligned (rows=(0, n)) and one at a time otherwise (rows=(1, n))."""
    g = torch.Generator().manual_seed(5)
    H, W, ps = 120, 140, 96            # 140 int16 = 280 B per count row: rows=(1, ..) is not 16-byte aligned
    d, sc = torch.rand(ps, ps, generator=g), torch.rand(ps, ps, generator=g)
    maps = [torch.rand(H, W, generator=g).cuda() for _ in range(4)] + [torch.ones(H, W, dtype=torch.int16, device="cuda")]
    ref = [m.cpu().clone() for m in maps]
    r0, r1, (c0, c1), (y0, x0) = 16, 80, cols, origin
    ops.accumulate_tile(d.cuda(), sc.cuda(), (r0, r1), (c0, c1), maps, y0, x0)
    sl = (slice(y0 + r0, y0 + r1), slice(x0 + c0, x0 + c1))
    dd, ss = d[r0:r1, c0:c1], sc[r0:r1, c0:c1]
    ref[0][sl] += dd; ref[1][sl] += dd * dd; ref[2][sl] += ss; ref[3][sl] += ss * ss; ref[4][sl] += 1
    for a, b in zip(maps, ref):
        assert torch.equal(a.cpu(), b)
    for rows in ((0, H), (1, H - 1)):
        fin = [m.clone() for m in maps]
        ops.finalize_map(fin, rows=rows)
        want = [m.clone() for m in ref]
        blk = slice(rows[0], rows[1])
        div = torch.zeros(H, W, dtype=torch.bool)
        div[blk] = want[4][blk] > 1
        cf = want[4][div].float()
        want[0][div] = want[0][div] / cf
        want[1][div] = torch.sqrt((want[1][div] - want[0][div] ** 2 * cf) / (cf - 1))
        want[2][div] = want[2][div] / cf
        want[3][div] = torch.sqrt((want[3][div] - want[2][div] ** 2 * cf) / (cf - 1))
        for a, b in zip(fin[:4], want[:4]):
            assert torch.allclose(a.cpu(), b, rtol=1e-6, atol=0, equal_nan=True)


@pytest.mark.parametrize("head_in", [16, 8])
@pytest.mark.parametrize("B,H,W", [(1, 64, 96), (2, 37, 53), (1, 300, 517)])
def test_head_tensor_core_matches_simt_and_oracle(head_in, B, H, W):
    """tcgen05 split-operand head (activations in TMEM) vs the fp32 SIMT head and the torch fp32 MLP."""
    from popcorn_b200 import weights
    g = torch.Generator().manual_seed(H + head_in)
    sd = po.random_state_dict(seed=3, head_in=head_in)
    feats = (torch.randn(B, head_in, H, W, generator=g) * 3.0)
    bu = torch.rand(B, 1, H, W, generator=g)
    ids = torch.randint(0, 5, (B, H, W), generator=g).to(torch.int32)
    ref_out = po.head_mlp(sd, feats.permute(0, 2, 3, 1).reshape(-1, head_in)).view(B, H, W, 2)[..., 0]
    ref_scale = F.relu(ref_out)
    ref_dens = ref_scale * bu[:, 0]
    outs = {}
    for tc in (False, True):
        pack = (weights.pack_head_tc(sd) if tc else weights.pack_head(sd)).cuda()
        sums = torch.zeros(5, dtype=torch.float64, device="cuda")
        dens, scale = ops.head_dense_forward(pack, feats.cuda(), bu.cuda(), ids.cuda(), None, sums, tc=tc)
        torch.cuda.synchronize()
        outs[tc] = (dens.cpu(), scale.cpu(), sums.cpu())
        err = ((scale.cpu() - ref_scale).abs() / ref_scale.abs().clamp(min=1e-3 * float(ref_scale.abs().max()))).max()
        assert float(err) < 1e-4, (tc, float(err))
        assert torch.allclose(dens.cpu(), ref_dens, rtol=1e-4, atol=1e-4 * float(ref_dens.abs().max()))
        ref_sums = torch.zeros(5, dtype=torch.float64).index_add_(0, ids.reshape(-1).long(), ref_dens.reshape(-1).double())
        assert torch.allclose(sums.cpu(), ref_sums, rtol=1e-5)
    assert torch.allclose(outs[True][0], outs[False][0], rtol=1e-4, atol=1e-4 * float(ref_dens.abs().max()))
