import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), d['gpu_launches']/d['steps'])
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1].get('share',0))[:3]:
    print('  ',k, v['calls']//d['steps'], v['ms_avg'], v.get('achieved'), v.get('unit'), v.get('frac'))
