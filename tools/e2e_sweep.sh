for cf in 157000 313000 625000; do
  timeout 300 python bench.py --config cfg3 --no-secondary --no-cpu-baseline --no-elbo-check --e2e-chunk-frames $cf > gpurun_out/e2e_$cf.json 2>gpurun_out/e2e_$cf.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/e2e_$cf.json').read().strip().splitlines()[-1])
print($cf, d['ms_per_step'], d['e2e'])
"
done
for cf in 512000 1024000 2048000; do
  timeout 300 python bench.py --config cfg2 --no-secondary --no-cpu-baseline --no-elbo-check --e2e-chunk-frames $cf > gpurun_out/e2e2_$cf.json 2>gpurun_out/e2e2_$cf.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/e2e2_$cf.json').read().strip().splitlines()[-1])
print($cf, d['ms_per_step'], d['e2e'])
"
done
