show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'headline', round(d['ms_per_step'],4), 'pre100h', round(d['legs']['preprocess_100h']['ms_per_step'],3), 'p2', round(d['legs']['stft_p2']['ms_per_step'],4))" $1 $2; }
cp ml_audio_inpainting_b200/lib/libaip_b200.so /tmp/orig.so
for r in 1 2; do
python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --gl-clips 0 > /tmp/base.json 2>/dev/null; show /tmp/base.json base
cp build/ab/skiptail.so ml_audio_inpainting_b200/lib/libaip_b200.so
python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --gl-clips 0 > /tmp/skip.json 2>/dev/null; show /tmp/skip.json skip
[ $r = 1 ] && python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
cp /tmp/orig.so ml_audio_inpainting_b200/lib/libaip_b200.so
done
