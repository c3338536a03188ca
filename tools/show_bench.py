import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
        print(f, round(d['value'],1), round(d['ms_per_step'],2), d['config']['solver_iters_last_step'], d['roofline'].get('amg_layout'), 'e2e', round(d['e2e']['value'],1))
        print('   ', d['roofline']['top_kernels_launches_ms_GBps'])
    except Exception as e:
        print(f, 'ERR', e)
