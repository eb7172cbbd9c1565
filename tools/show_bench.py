"""Print the interesting parts of one or more bench.py JSON lines (files given on the command line)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    print(f"{f}: N={d['n_gpus']} value {d['value']:.2f} ms  e2e {d['e2e']['value']:.2f} ms  arms={d.get('arms_bit_identical')} "
          f"r2r={d.get('run_to_run_bit_identical')} ranks={d.get('ranks_bit_identical')} launches={d['gpu_launches']}")
    sp = d.get("parity_spot")
    if sp:
        print(f"   spot J {sp['max_scaled_J']:.1e} K {sp['max_scaled_K']:.1e} ({sp['seconds']:.1f} s) ok={sp['ok']}")

    def kern(ks, ind="   "):
        for k, v in ks.items():
            print(ind, k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if not isinstance(b, (dict, str))})
            for ph, pv in (v.get("int8_arm") or {}).items():
                if isinstance(pv, dict):
                    print(ind, "    i8", ph, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in pv.items() if not isinstance(b, (dict, str))})
                else:
                    print(ind, "    i8", ph, pv)

    kern(d.get("kernels", {}))
    r = d.get("roofline")
    if r:
        print(f"   roofline [{r.get('kernel', '')[:40]}] {r['achieved']:.2f} / {r['peak']:.2f} {r.get('unit')} = {r['frac']:.3f}")
    r = d.get("roofline_fp64")
    if r:
        print(f"   roofline_fp64 {r['achieved']:.2f} / {r['peak']:.2f} = {r['frac']:.3f}")
    for k, v in d.get("workloads", {}).items():
        s2 = v["parity_spot"]
        print(f"   WL {k}: value {v['value_ms']:.2f} e2e {v['e2e_ms']:.2f} spot "
              f"{(s2['max_scaled_J'], s2['max_scaled_K']) if s2 else None} arms={v['arms_bit_identical']} r2r={v['run_to_run_bit_identical']}")
        kern(v["kernels"], "      ")
    cb = d.get("cpu_baseline")
    print("   cpu", cb and round(cb["value"], 1), cb and cb["cores"], " clocks", d.get("clocks"))
    ab = d.get("fp64_arms")
    if ab:
        print("   fp64 arms:", {a: (round(b, 3) if isinstance(b, float) else b) for a, b in ab.items() if a != "what"})
