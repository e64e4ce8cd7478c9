"""Numerics of the TMA-fed tcgen05 GEMM kernel (csrc/cuda/tc_gemm.cu) against plain fp32 PyTorch.

Three operand layouts are exercised: K-major x K-major (forward), K-major x MN-major (dA) and MN-major x MN-major
with split-K fp32 reductions (dW); the batch-dependent extent comes from a device counter like in the engine."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _counters(dev, idx, val):
  c = torch.zeros(16, dtype=torch.int32, device=dev)
  c[idx] = val
  return c


def _rand(shape, dev, scale=1.0, seed=0):
  g = torch.Generator(device=dev)
  g.manual_seed(seed)
  return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize('cap,T,K,N,relu,with_bias', [
  (1000, 777, 512, 256, True, True),
  (1024, 1024, 512, 64, False, True),
  (40000, 33333, 256, 256, True, False),
  (300, 1, 128, 128, False, True),
  (512, 0, 512, 256, True, True),
])
def test_tc_forward_matches_fp32(native, cap, T, K, N, relu, with_bias):
  dev = torch.device('cuda', 0)
  A = _rand((cap, K), dev, seed=1)
  W = _rand((N, K), dev, 0.05, seed=2)
  b = _rand((N,), dev, seed=3) if with_bias else None
  Z = torch.full((cap, N), 7.0, dtype=torch.bfloat16, device=dev)
  c = _counters(dev, 2, T)
  pl = native.TcGemm(0)
  pl.add_forward(A, W, b, relu, Z, c, 2)
  pl.run()
  torch.cuda.synchronize()
  ref = A[:T].float() @ W.float().t()
  if with_bias:
    ref = ref + b.float()
  if relu:
    ref = ref.relu()
  if T > 0:
    assert torch.allclose(Z[:T].float(), ref, rtol=2e-2, atol=2e-2), float((Z[:T].float() - ref).abs().max())
  # rows of untouched tiles keep their old contents (only tiles below ceil(T/128) are written)
  first_untouched = (T + 127) // 128 * 128
  if first_untouched < cap:
    assert bool((Z[first_untouched:] == 7.0).all())
  # the counter is re-read on every launch: shrink the batch and run the SAME plan again
  if T > 200:
    c[2] = 130
    Z.fill_(7.0)
    pl.run()
    torch.cuda.synchronize()
    assert torch.allclose(Z[:130].float(), ref[:130], rtol=2e-2, atol=2e-2)
    assert bool((Z[256:] == 7.0).all())


@pytest.mark.parametrize('cap,T,Kd,N', [(1200, 1000, 256, 512), (1024, 1024, 64, 512), (20000, 12345, 256, 256)])
def test_tc_dgrad_matches_fp32(native, cap, T, Kd, N):
  dev = torch.device('cuda', 0)
  dPre = _rand((cap, Kd), dev, seed=4)
  W = _rand((Kd, N), dev, 0.05, seed=5)
  dA = torch.zeros(cap, N, dtype=torch.bfloat16, device=dev)
  c = _counters(dev, 1, T)
  pl = native.TcGemm(0)
  pl.add_dgrad(dPre, W, dA, c, 1)
  pl.run()
  torch.cuda.synchronize()
  ref = dPre[:T].float() @ W.float()
  assert torch.allclose(dA[:T].float(), ref, rtol=2e-2, atol=2e-2), float((dA[:T].float() - ref).abs().max())


@pytest.mark.parametrize('cap,T,M,N', [(1200, 1000, 256, 512), (1024, 1024, 64, 512), (50000, 45678, 256, 256),
                                       (4096, 63, 256, 512), (4096, 0, 256, 512)])
def test_tc_wgrad_split_k_matches_fp32(native, cap, T, M, N):
  dev = torch.device('cuda', 0)
  dPre = _rand((cap, M), dev, 0.1, seed=6)
  dPre[T:] = 0                      # engine invariant: rows beyond the batch are zero up to the capacity
  A = _rand((cap, N), dev, seed=7)  # rows beyond T hold stale (finite) data
  gW = torch.zeros(M, N, dtype=torch.float32, device=dev)
  c = _counters(dev, 3, T)
  pl = native.TcGemm(0)
  pl.add_wgrad(dPre, A, gW, c, 3)
  pl.run()
  torch.cuda.synchronize()
  ref = dPre[:T].float().t() @ A[:T].float()
  scale = max(1.0, float(ref.abs().max()))
  assert float((gW - ref).abs().max()) <= 2e-3 * scale, (float((gW - ref).abs().max()), scale)
  # red-add semantics: a second run accumulates
  pl.run()
  torch.cuda.synchronize()
  assert float((gW - 2 * ref).abs().max()) <= 4e-3 * scale


def test_tc_wgrad_and_dgrad_share_one_launch(native):
  dev = torch.device('cuda', 0)
  cap, T, n_out, two_d = 12000, 9999, 256, 512
  dPre = _rand((cap, n_out), dev, 0.1, seed=8)
  dPre[T:] = 0
  A = _rand((cap, two_d), dev, seed=9)
  W = _rand((n_out, two_d), dev, 0.05, seed=10)
  gW = torch.zeros(n_out, two_d, dtype=torch.float32, device=dev)
  dA = torch.zeros(cap, two_d, dtype=torch.bfloat16, device=dev)
  c = _counters(dev, 2, T)
  pl = native.TcGemm(0)
  pl.add_wgrad(dPre, A, gW, c, 2)
  pl.add_dgrad(dPre, W, dA, c, 2)
  pl.run()
  torch.cuda.synchronize()
  ref_w = dPre[:T].float().t() @ A[:T].float()
  ref_a = dPre[:T].float() @ W.float()
  assert float((gW - ref_w).abs().max()) <= 2e-3 * max(1.0, float(ref_w.abs().max()))
  assert torch.allclose(dA[:T].float(), ref_a, rtol=2e-2, atol=2e-2)


def test_tc_gemm_inside_cuda_graph(native):
  dev = torch.device('cuda', 0)
  A = _rand((2048, 512), dev, seed=11)
  W = _rand((256, 512), dev, 0.05, seed=12)
  Z = torch.zeros(2048, 256, dtype=torch.bfloat16, device=dev)
  c = _counters(dev, 1, 1500)
  pl = native.TcGemm(0)
  pl.add_forward(A, W, None, False, Z, c, 1)
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    pl.run()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
      pl.run()
  torch.cuda.synchronize()
  Z.zero_()
  c[1] = 2048
  g.replay()
  torch.cuda.synchronize()
  ref = A.float() @ W.float().t()
  assert torch.allclose(Z.float(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize('M,T,K,N,relu', [(512, 512, 256, 256, False), (1000, 900, 512, 128, True), (4096, 4000, 1024, 512, False)])
def test_tc_block_scaled_mxfp8_gemm_matches_dequantised_fp32(native, M, T, K, N, relu):
  """tcgen05.mma.kind::mxf8f6f4.block_scale: e4m3 operands with UE8M0 scales per 32 K-elements; the result must
  equal the fp32 product of the DE-QUANTISED operands (the quantisation error itself is not part of the kernel)."""
  from graphlearn_for_pytorch_b200.data import dequantize_mxfp8, pack_mx_scale_blocks, quantize_mxfp8, quantize_mxfp8_parts
  dev = torch.device('cuda', 0)
  g = torch.Generator(device=dev); g.manual_seed(5)
  A = torch.randn(M, K, device=dev, generator=g) * (0.1 + 4 * torch.rand(M, 1, device=dev, generator=g))
  W = torch.randn(N, K, device=dev, generator=g) * 0.05
  b = _rand((N,), dev, seed=9)
  Aq, Asf = quantize_mxfp8_parts(A)
  Wq, Wsf = quantize_mxfp8_parts(W)
  Ad = dequantize_mxfp8(quantize_mxfp8(A), K)
  Wd = dequantize_mxfp8(quantize_mxfp8(W), K)
  Z = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=dev)
  c = _counters(dev, 1, T)
  pl = native.TcGemmMx(0, Aq, pack_mx_scale_blocks(Asf), Wq, pack_mx_scale_blocks(Wsf), b, relu, Z, c, 1)
  pl.run()
  torch.cuda.synchronize()
  ref = Ad[:T] @ Wd.t() + b.float()
  if relu:
    ref = ref.relu()
  err = float((Z[:T].float() - ref).abs().max())
  scale = max(1.0, float(ref.abs().max()))
  assert err <= 1.5e-2 * scale, (err, scale)
