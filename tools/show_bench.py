import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, 'ms/step', round(d["ms_per_step"],3), 'frac', round(d["roofline"]["step"]["frac_of_burst_peak"],4), 'e2e', round(d["e2e"]["ms_per_step"],3), 'launches', d["gpu_launches"], d["clocks"]["reasons"])
    for k,v in (d["roofline"]["kernels"].items() if "-v" in sys.argv else []): print('   ', k[:28], round(v["ms"],3), round(v["tflops"]))
