#!/bin/bash
# quick A/B of environment switches on the bench step: prints value and the big families
for v in "$@"; do
  env $v python bench.py --steps 5 --warmup 3 --corpus 0 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=l['families']
print('$v', 'value %.1f ms/step %.1f' % (l['value'], l['ms_per_step']), {k: round(f[k]['ms_per_step'],1) for k in ('flash_attn','ffn_fused','gemm_tap<256>','gemm_tap<128>','gemm_tap<64>')}, 'qkv %.1f conv1 %.1f conv2 %.1f' % tuple(l['gemm256_by_epilogue'][k]['ms_per_step'] for k in ('qkv_split', 'conv+ln+mish+temb', 'conv+ln+mish+res+ln_emit')))
"
done
