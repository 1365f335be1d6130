import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        line = line.strip()
        if line.startswith('{'):
            d = json.loads(line)
            print(f, 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 2),
                  'launches', d.get('gpu_launches'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
            print('  ', d.get('breakdown_ms_per_step'))
            print('  ', {k: round(v['frac'], 3) for k, v in d.get('kernels', {}).items()})
            if 'trajectory' in d: print('  ', d['trajectory'])
