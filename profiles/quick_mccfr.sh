# headline kernel only: parity tests, then the bench line's value / kernel time
python -m pytest tests/test_gpu_solver.py tests/test_gpu_dropin.py tests/test_gpu_multideal.py -x -q 2>&1 | tail -2
for TRAV in 340992 454656; do
python bench.py --steps 10 --warmup 3 --no-cpu --md-deals 0 --full-games 0 --trav $TRAV 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('trav', d['config']['traversals_per_step'], 'G upd/s', d['value']/1e9, 'ms/step', d['ms_per_step'], 'kernel ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])"
done
