"""Reads a bench.py JSON line on stdin and prints value, ms/step, launches/step and the per-kernel table (top 3 by share):
   python bench.py --robots 64 --graphs 128 --no-e2e --no-cpu-baseline --no-other-configs | python benchmarks/kernel_table.py"""
import json,sys
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)     # `| head` closes the pipe early
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), d['gpu_launches']/d['steps'])
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1].get('share',0))[:3]:
    print('  ',k, v['calls']//d['steps'], v['ms_avg'], v.get('achieved'), v.get('unit'), v.get('frac'))
