import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('value %.3e %s  ms/step %.2f  n_gpus %d' % (d['value'], d['unit'], d['ms_per_step'], d['n_gpus']))
        r = d.get('roofline') or {}
        print('roofline', {k: r.get(k) for k in ('achieved', 'peak', 'frac', 'kernel_ms', 'fp64_equivalent_tflops')})
        print('kinship', d.get('kinship'))
        print('glm', d.get('glm'))
        e = d.get('e2e') or {}
        print('e2e', e.get('value'), e.get('ms_per_step'))
        print('cpu', d.get('cpu_baseline'))
        print('clocks', d.get('clocks'))
        print('setup', d.get('setup_ms'))
    elif 'exit' in l or 'Error' in l or 'error' in l:
        print(l.rstrip())
