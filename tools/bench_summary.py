#!/usr/bin/env python
"""One line per bench JSON file: step time, kernel split, e2e, overflow bins."""
import json
import sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        km = d['roofline']['kernel_ms']
        print('%-28s ms/step %7.1f  value %.3g  e2e %7.1f | k1 %6.1f k2a %6.1f k2b %6.1f k3 %6.1f | ovf %d retries %d bins %d pair %.4f' % (
            f.split('/')[-1], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], km['k1_superkmer_partition'], km['k2a_fine_split'],
            km['k2b_bucket_hash_count'], km["k3_partition_id_sort"], d["overflow_bins"], d['retries'], d['bins'], d['roofline']['pair']['frac']))
        if d.get('e2e_api'):
            print('    e2e_api', json.dumps(d['e2e_api'])[:400])
        if d.get('cpu_baseline'):
            print('    cpu_baseline', json.dumps(d['cpu_baseline'])[:300])
        print('    partitions', d['config'].get('nb_partitions'), d['config'].get('repartitor'), 'invariants', d.get('invariants', {}).get('all'))
    except Exception as e:
        print(f, 'ERR', e)
