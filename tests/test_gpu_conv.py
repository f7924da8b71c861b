"""Parity of the tcgen05 implicit-GEMM conv (skb_conv2d_bf16, through the C ABI) against the CPU
oracle arithmetic of ConvolutionBlock (oracle/model.py: conv_block), on bf16-representable inputs.

Tolerances (relative to max |reference|): fp32-accumulate mode = bf16 operands, fp32 accumulation,
fp32 store: <= 1e-3 (north_star bound); bf16 store adds one rounding: <= 6e-3 (stated bf16 bound
for a single layer = 2^-8 half-ulp * margin)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16r, randn, rel_err

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-3
TOL_BF16 = 6e-3


def _ref(x, w, b, stride, act, res=None):
    k = w.shape[-1]
    y = F.conv2d(x, w, b, stride, k // 2)
    if act == 1:
        y = F.silu(y)
    elif act == 2:
        y = F.relu(y)
    if res is not None:
        y = y + res
    return y


def _run(n, h, w, cin, cout, k, stride, act=1, residual=False, up=False, out_f32=True, slice_in=0, slice_out=0, seed=0):
    from skyeye import engine as E
    x = bf16r(randn(("cx", seed, n, h, w, cin), (n, cin, h, w)))
    wt = bf16r(randn(("cw", seed, cout, cin, k), (cout, cin, k, k), (2.0 / (cin * k * k)) ** 0.5))
    b = randn(("cb", seed, cout), (cout,), 0.1)
    ho, wo = h // stride, w // stride
    res = bf16r(randn(("cr", seed), (n, cout, ho, wo))) if residual else None
    ref = _ref(x, wt, b, stride, act, res)
    if up:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
    # input as a channel slice of a wider buffer
    xb = torch.zeros((n, h, w, cin + slice_in), dtype=torch.bfloat16, device="cuda")
    xb[..., slice_in:] = x.permute(0, 2, 3, 1).to(torch.bfloat16).cuda()
    xv = E.View(xb, slice_in, cin)
    pw = E.PackedConv(wt, b)
    c8 = (cout + 7) // 8 * 8
    u = 2 if up else 1
    yb = torch.full((n, ho * u, wo * u, c8 + slice_out), 7.0, dtype=torch.float32 if out_f32 else torch.bfloat16, device="cuda")
    yv = E.View(yb, slice_out, c8)
    rv = None
    if residual:
        rb = res.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
        rv = E.View(rb)
    E.conv2d(xv, pw, yv, stride, act, rv, up)
    torch.cuda.synchronize()
    got = yv.torch()[..., :cout].permute(0, 3, 1, 2).float().cpu()
    if slice_out:
        assert bool((yb[..., :slice_out] == 7.0).all()), "conv wrote outside its channel slice"
    return rel_err(got, ref)


CASES = [
    # n, h, w, cin, cout, k, stride
    (2, 16, 16, 64, 64, 1, 1),
    (2, 16, 16, 64, 64, 3, 1),
    (1, 32, 32, 128, 128, 3, 1),
    (2, 16, 16, 64, 128, 3, 2),
    (1, 20, 20, 256, 256, 3, 1),      # ragged spatial tiles
    (3, 10, 10, 512, 512, 1, 1),      # tile spans several images
    (1, 40, 40, 128, 256, 3, 2),
    (2, 8, 8, 1024, 512, 1, 1),
    (2, 16, 16, 32, 64, 3, 2),        # BK = 32 path (skyeye_s stem)
    (2, 16, 16, 32, 32, 3, 1),        # BN = 32
    (2, 16, 16, 96, 64, 1, 1),        # Cin % 64 != 0 -> BK = 32
    (1, 16, 16, 256, 45, 1, 1),       # detection head: Cout 45 -> padded, fp32 out
    (1, 6, 6, 2048, 1024, 1, 1),      # SPP cv2 shape class
    (1, 24, 40, 64, 192, 1, 1),       # non-square, Cout multiple of 64 only
    (2, 32, 48, 16, 64, 3, 1),        # BK = 16 path (Focus conv: 12 channels padded to 16, SWIZZLE_32B)
    (1, 16, 16, 16, 32, 3, 1),        # BK = 16, BN = 32 (skyeye_s stem)
    (1, 16, 16, 48, 64, 1, 1),        # Cin % 32 != 0 -> BK = 16, three K chunks
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_fp32_accumulate_mode(case):
    err = _run(*case, out_f32=True)
    assert err < TOL_F32, err


@pytest.mark.parametrize("case", CASES[:8], ids=lambda c: "x".join(map(str, c)))
def test_conv_bf16_store(case):
    err = _run(*case, out_f32=False)
    assert err < TOL_BF16, err


def test_conv_halo_kernel_residual_and_slices():
    # 3x3 stride-1 layers whose map is covered by 16x16 tiles run the halo-tile kernel (conv_gemm.cu)
    assert _run(2, 32, 32, 128, 128, 3, 1, residual=True, out_f32=False) < TOL_BF16      # BN = 128: one staging slot
    assert _run(1, 32, 48, 64, 64, 3, 1, residual=True, out_f32=False) < TOL_BF16        # BN = 64: two staging slots
    assert _run(1, 16, 32, 64, 64, 3, 1, slice_in=64, slice_out=64, out_f32=False) < TOL_BF16
    assert _run(3, 16, 16, 256, 256, 3, 1, residual=True, out_f32=False) < TOL_BF16      # 4 chunks x 2 channel blocks
    assert _run(1, 30, 46, 64, 128, 3, 1, out_f32=False) < TOL_BF16                      # ragged tiles in both directions
    assert _run(2, 32, 16, 192, 64, 3, 1, out_f32=False, act=2) < TOL_BF16               # three chunks, ReLU
    assert _run(1, 48, 32, 128, 256, 3, 1, out_f32=False, act=0) < TOL_BF16              # two channel blocks, linear


def test_conv_residual_after_activation():
    # the residual path stores bf16 (bottlenecks, transformer residuals); fp32 + residual is not on the path
    assert _run(2, 16, 16, 128, 128, 3, 1, residual=True, out_f32=False) < TOL_BF16
    assert _run(1, 20, 20, 256, 256, 3, 1, residual=True, out_f32=False) < TOL_BF16
    assert _run(3, 10, 10, 64, 32, 1, 1, residual=True, out_f32=False) < TOL_BF16   # 64-byte rows (BN = 32)


def test_conv_relu_and_linear():
    assert _run(1, 16, 16, 64, 256, 1, 1, act=2) < TOL_F32
    assert _run(1, 16, 16, 64, 256, 1, 1, act=0) < TOL_F32


def test_conv_fused_upsample():
    assert _run(2, 8, 8, 128, 64, 1, 1, up=True) < TOL_F32
    assert _run(2, 8, 8, 128, 64, 1, 1, up=True, out_f32=False) < TOL_BF16


def test_conv_reads_and_writes_channel_slices():
    assert _run(2, 16, 16, 64, 64, 3, 1, slice_in=64, slice_out=64) < TOL_F32
    assert _run(2, 16, 16, 64, 64, 3, 2, slice_in=32, slice_out=8, out_f32=False) < TOL_BF16


def test_conv_inplace_bottleneck_update():
    """y = y + silu(conv3x3(t)) with residual aliasing the output slice (CSP bottleneck chain)."""
    from skyeye import engine as E
    n, h, w, c = 2, 16, 16, 64
    t = bf16r(randn(("ipt",), (n, c, h, w)))
    y0 = bf16r(randn(("ipy",), (n, c, h, w)))
    wt = bf16r(randn(("ipw",), (c, c, 3, 3), 0.05))
    b = randn(("ipb",), (c,), 0.1)
    ref = bf16r(y0 + F.silu(F.conv2d(t, wt, b, 1, 1)))
    cat = torch.zeros((n, h, w, 2 * c), dtype=torch.bfloat16, device="cuda")
    cat[..., :c] = y0.permute(0, 2, 3, 1).to(torch.bfloat16).cuda()
    yv = E.View(cat, 0, c)
    E.conv2d(E.from_nchw(t.cuda()), E.PackedConv(wt, b), yv, 1, 1, yv)
    torch.cuda.synchronize()
    assert rel_err(yv.nchw(), ref) < TOL_BF16
    assert bool((cat[..., c:] == 0).all())


def test_conv_large_layer_matches_sampled_reference():
    """BASELINE-size layer class (128->128 3x3 @160x160, B=4): compare a sample of output pixels."""
    from skyeye import engine as E
    n, h, w, c = 4, 160, 160, 128
    x = bf16r(randn(("lx",), (n, c, h, w)))
    wt = bf16r(randn(("lw",), (c, c, 3, 3), (2.0 / (9 * c)) ** 0.5))
    b = randn(("lb",), (c,), 0.1)
    y = E.new_buffer(n, h, w, c, torch.float32)
    E.conv2d(E.from_nchw(x.cuda()), E.PackedConv(wt, b), y, 1, 1)
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(x[:, :, 60:100, :], wt, b, 1, 1))[:, :, 1:-1, :]  # rows 61..98, full width
    got = y.nchw()[:, :, 61:99, :].cpu()
    assert rel_err(got, ref) < TOL_F32


def test_conv_rejects_bad_arguments_loudly():
    from skyeye import engine as E
    x = E.new_buffer(1, 8, 8, 40)  # Cin not a multiple of 16
    with pytest.raises(AssertionError):
        E.conv2d(x, E.PackedConv(torch.zeros(64, 64, 1, 1), None), E.new_buffer(1, 8, 8, 64))
    pw = E.PackedConv(torch.zeros(64, 40, 1, 1), None)
    with pytest.raises(RuntimeError, match="multiple of 16"):
        E.conv2d(x, pw, E.new_buffer(1, 8, 8, 64))
