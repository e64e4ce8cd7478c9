"""MXFP8 feature storage (OCP microscaling format): e4m3 elements with one power-of-two UE8M0 scale per block of 32.

The reference stores features as fp32 / fp16 / bf16 tensors (python/data/feature.py; csrc/cuda/unified_tensor.cu:112-130
registers the dtypes).  On B200 the feature gather of layer 1 is bound by bytes moved (local HBM and, for partitioned
tables, NVLink), so halving the row size is worth more than any kernel tweak: a d-wide row becomes

    [ d bytes e4m3 | d/32 bytes UE8M0 scales | zero padding to a multiple of 16 bytes ]

(d = 128: 144 bytes instead of 256).  The fused layer-1 kernel (csrc/cuda/sage_tc.cu, `FP8` loader) and
`RowTableHandle.gather_mxfp8` de-quantise in registers; everything downstream still sees bf16 / fp32.
"""
import torch

BLOCK = 32
E4M3_MAX = 448.0


def mxfp8_row_bytes(d: int) -> int:
  return (d + d // BLOCK + 15) // 16 * 16


def quantize_mxfp8(x: torch.Tensor) -> torch.Tensor:
  """[N, d] float/bf16 (d a multiple of 32)  ->  uint8 [N, mxfp8_row_bytes(d)] packed rows."""
  assert x.dim() == 2 and x.shape[1] % BLOCK == 0
  n, d = x.shape
  xf = x.float().view(n, d // BLOCK, BLOCK)
  amax = xf.abs().amax(dim=2)
  # smallest power of two s with amax / s <= 448
  e = torch.ceil(torch.log2(torch.clamp(amax, min=1e-30) / E4M3_MAX)).clamp(-126, 127)
  e = torch.where(amax > 0, e, torch.full_like(e, -126))
  scale = torch.exp2(e).unsqueeze(2)
  q = (xf / scale).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).view(n, d)
  out = torch.zeros(n, mxfp8_row_bytes(d), dtype=torch.uint8, device=x.device)
  out[:, :d] = q.view(torch.uint8)
  out[:, d:d + d // BLOCK] = (e + 127).to(torch.uint8)
  return out


def dequantize_mxfp8(rows: torch.Tensor, d: int) -> torch.Tensor:
  """Inverse of `quantize_mxfp8` (fp32 [N, d]); the reference semantics for the device de-quantisers."""
  n = rows.shape[0]
  q = rows[:, :d].contiguous().view(torch.float8_e4m3fn).float().view(n, d // BLOCK, BLOCK)
  e = rows[:, d:d + d // BLOCK].float() - 127.0
  return (q * torch.exp2(e).unsqueeze(2)).view(n, d)


def quantize_mxfp8_parts(x: torch.Tensor):
  """[N, d] -> (e4m3 bytes uint8 [N, d], UE8M0 scale bytes uint8 [N, d / 32]): the operands of a block-scaled
  tcgen05 GEMM (`TcGemmMx`); the scale bytes still have to be packed with `pack_mx_scale_blocks`."""
  rows = quantize_mxfp8(x)
  d = x.shape[1]
  return rows[:, :d].contiguous(), rows[:, d:d + d // BLOCK].contiguous()


def pack_mx_scale_blocks(sf: torch.Tensor) -> torch.Tensor:
  """UE8M0 scale bytes [rows, K / 32] -> the tensor core's block layout: uint8 [ceil(rows/128), K/128, 512], where
  byte (r % 32) * 16 + (r // 32) * 4 + k of block (mb, kb) is the scale of row mb * 128 + r, K-group kb * 4 + k
  (CUTLASS Sm1xxBlockScaledBasicChunk; rows are padded with the neutral scale 2^0)."""
  rows, g = sf.shape
  assert g % 4 == 0, 'K must be a multiple of 128'
  mb, kb = (rows + 127) // 128, g // 4
  pad = torch.full((mb * 128, g), 127, dtype=torch.uint8, device=sf.device)
  pad[:rows] = sf
  v = pad.view(mb, 4, 32, kb, 4)                 # [mb, r // 32, r % 32, kb, k]
  return v.permute(0, 3, 2, 1, 4).contiguous().view(mb, kb, 512)
