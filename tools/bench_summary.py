import json, sys
d = json.load(open(sys.argv[1]))
print('%.2f Gpts/s  %.0f fps  %.3f ms/step | e2e %.2f Gpts/s | p50 %.3f ms | pipe frac %.4f | launches %d' % (
    d['value']/1e9, d['frames_per_sec'], d['ms_per_step'], d['e2e']['value']/1e9, (d.get('latency') or {}).get('p50_ms', 0),
    d['pipeline_roofline']['frac'], d['gpu_launches']))
print(d['stage_ms_per_step'])
r = d.get('roofline') or {}
print('roofline', r.get('kernel'), r.get('achieved'), r.get('frac'))
for k, v in list(d['kernels'].items())[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print('  %-20s %9.1f us %4d  %.3f' % (k, v['total_us'], v['launches'], v['share']))
if d.get('cpu_baseline'): print('cpu', d['cpu_baseline']['value']/1e6, 'Mpts/s 1 core;', d['cpu_baseline']['all_cores'])
