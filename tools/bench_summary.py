import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        km=d["roofline"]["kernel_ms"]
        print(f, "%.3f G/s frac %.3f"%(d["value"]/1e9, d["step_hbm_frac"]), {k:round(v*1e3,1) for k,v in km.items()})
    except Exception as e:
        print(f, "ERR", e)
