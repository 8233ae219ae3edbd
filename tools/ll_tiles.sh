for cfg in 512,64,8 1024,64,8 2048,128,8 4096,128,8 4096,128,2 2048,128,6; do
for rows in 2 4; do
TAME_LLMSE=dfma TAME_LL_ROWS=$rows python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith(chr(123))][-1]
print('$cfg rows $rows', round(d['ms_per_step'],3), round(d['roofline']['kernels_ms_per_step']['k_llmse'],4))"
done; done
