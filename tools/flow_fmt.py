import sys, json
for l in sys.stdin:
    if l.startswith('RESULT '):
        o = json.loads(l[7:])
        for p in o['parity']: print('parity', p['n'], p['k'], 'max', p['relerr'], p['nonfinite_match'], ['%.1e' % e for e in p.get('each', [])])
        for t in o['timing']: print('timing n=%d k=%d %.2f ms %.4g evals/s %.2f TF checksum %.12g' % (t['n'], t['k'], t['ms'], t['evals_per_s'], t['tflops'], t['checksum']))
    else: print(l, end='')
