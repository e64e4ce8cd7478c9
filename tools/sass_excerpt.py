"""Write profiles/sass_excerpts_tcgen05_kernels.txt: verbatim SASS lines (tensor-core, TMEM, TMA, PDL instructions) of
the tcgen05 kernels in the built extension, plus a per-kernel mnemonic census.

  cuobjdump -sass graphlearn_for_pytorch_b200/_ext/glt_b200_C.so > /tmp/all.sass && python tools/sass_excerpt.py /tmp/all.sass
"""
import re
import sys

KERNELS = (('k_sage_fused3ILi2ELb0', 'k_sage_fused3<2,false>  (fused gather + mean + [mean|self].W^T + bias + ReLU, bf16 rows)'),
           ('k_sage_fused3ILi2ELb1', 'k_sage_fused3<2,true>   (same, MXFP8 feature rows de-quantised in the loaders)'),
           ('k_tc_gemm', 'k_tc_gemm  (TMA-fed forward / dA / split-K dW GEMMs)'))
CENSUS = re.compile(r'UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|SYNCS|UTMACCTL|UTMACMDFLUSH|ACQBULK|PREEXIT|REDG|F2FP|'
                    r'FENCE|MEMBAR|UTCCP')
SHOW = re.compile(r'UTC[A-Z]*MMA|LDTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|ACQBULK|PREEXIT|REDG|UTMACCTL|UTMACMDFLUSH')
OP = re.compile(r'\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Za-z0-9_]+)*)')


def main(path, out_path='profiles/sass_excerpts_tcgen05_kernels.txt'):
  lines = open(path).read().split('\n')
  starts = [(i, l) for i, l in enumerate(lines) if 'Function :' in l]
  out = ['# SASS excerpts of the tcgen05 kernels (cuobjdump -sass of graphlearn_for_pytorch_b200/_ext/glt_b200_C.so, sm_100a)',
         '# Every instruction line is copied verbatim from the listing; "..." marks elided instructions.  PTX -> SASS:',
         '#   tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, cp.async.bulk.tensor -> UTMALDG / UTMASTG,',
         '#   cp.async.bulk -> UBLKCP, mbarrier.* -> SYNCS.* (census only), griddepcontrol.wait -> ACQBULK,',
         '#   griddepcontrol.launch_dependents -> PREEXIT, red.global.add.v4.f32 -> REDG.E.ADD.F32x4, prefetch.tensormap -> UTMACCTL.PF']
  for key, title in KERNELS:
    for k, (i, l) in enumerate(starts):
      if key in l:
        body = lines[i:(starts[k + 1][0] if k + 1 < len(starts) else len(lines))]
        break
    else:
      continue
    out += ['', f'## {title}', body[0].strip()]
    census = {}
    for l in body:
      m = OP.search(l)
      if m and CENSUS.search(m.group(1)):
        census[m.group(1)] = census.get(m.group(1), 0) + 1
    out.append('census: ' + ', '.join(f'{k} x{v}' for k, v in sorted(census.items())))
    last = -10
    for idx, l in enumerate(body):
      m = OP.search(l)
      if m and SHOW.search(m.group(1)):
        if idx - last > 1:
          out.append('        ...')
        out.append(l.rstrip()[:150])
        last = idx
  open(out_path, 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
  main(*sys.argv[1:])
