import json, sys
for f in sys.argv[1:]:
    try:
        l = open(f).read().strip().splitlines()[-1]
        d = json.loads(l)
        print(f, "value %.4g e2e %.4g ms %.1f kernel_ms %.1f frac %.3f peak %.1f maxerr %.2e iters %.2f clocks %s" % (
            d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"],
            d["roofline"]["peak"], d["max_abs_err_vs_cpu_port"], d["config"]["mean_iterations"], d["clocks"]))
    except Exception as e:
        print(f, "ERR", open(f).read()[-400:])
