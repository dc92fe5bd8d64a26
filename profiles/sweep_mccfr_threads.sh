for T in 640 768; do
  sed -i "s/constexpr int MCCFR_THREADS = [0-9]*;/constexpr int MCCFR_THREADS = $T;/" scopa_b200/csrc/ms_solver.cu
  python scopa_b200/_build.py > /dev/null 2>&1
  B=$((148*T*3))
  python bench.py --steps 10 --warmup 3 --no-cpu --sd-trav 256 --games 1000000 --step-states 100000 --trav $B 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('T=$T', d['value']/1e9, 'G upd/s', d['ms_per_step'], 'rollout', d['env']['value']/1e9, d['env']['ms_per_step'])"
done
