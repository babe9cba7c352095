python bench.py --steps 4 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('plain', d['roofline']['kernel_ms'], d['fft_plan'])"
TA_B200_K1F_PROFILE=1 python bench.py --steps 2 --warmup 1 2>gpurun_out/exp_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('profiled', d['roofline']['kernel_ms'], d['fft_plan'])"; grep -A40 "k1f profile" gpurun_out/exp_err.txt | tail -32
