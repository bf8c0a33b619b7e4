#!/bin/bash
# sweep decode kernel launch parameters: group rows, stages, CTAs/SM, warps/CTA
for cfg in "32 2 2 4" "24 2 3 4" "16 2 2 8" "32 2 1 9" "16 3 2 5" "32 3 1 6" "12 2 3 8" "20 2 2 6"; do
  set -- $cfg
  MGD_DECODE_GROUP_ROWS=$1 MGD_DECODE_STAGES=$2 MGD_DECODE_CTAS_PER_SM=$3 MGD_DECODE_WARPS=$4 python bench.py --steps 5 --warmup 2 --batch 2048 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('rows/stages/ctas/warps $cfg', 'decode ms', round(k['decode_compact']['avg_launch_ms'],4), 'GB/s', round(k['decode_compact']['achieved_gbs']), 'fill', round(k['encode_fill']['achieved_gbs']), 'nms ms', round(k['nms']['avg_launch_ms'],4), 'img/s', round(d['value']))"
done
