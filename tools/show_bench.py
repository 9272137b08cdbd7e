"""Prints bench.py JSON lines side by side (A/B runs of the round scripts).

    python tools/show_bench.py [-v] FILE...      # FILE: a JSON file, or a log / .json with one bench line per text line
"""
import json
import sys


def lines_of(path):
    with open(path) as f:
        text = f.read()
    try:
        d = json.loads(text)
        return d if isinstance(d, list) else [d]
    except ValueError:
        return [json.loads(ln) for ln in text.splitlines() if ln.startswith("{") and '"metric"' in ln]


def main():
    verbose = "-v" in sys.argv
    for path in [a for a in sys.argv[1:] if a != "-v"]:
        try:
            recs = lines_of(path)
        except Exception as e:
            print(path, "ERR", e)
            continue
        for d in recs:
            knobs = d.get("config", {}).get("knobs", {})
            e = d.get("e2e", {})
            print(f'{path}: gpus {d.get("n_gpus")}  ms/step {d["ms_per_step"]:.3f}  value {d["value"] / 1e6:.2f} M/s  '
                  f'step frac {d["roofline"]["step"]["frac_of_burst_peak"]:.3f}  e2e {e.get("ms_per_step", float("nan")):.3f} ms '
                  f'({e.get("mode", "")[:9]})  launches {d.get("gpu_launches")}  clocks {d["clocks"].get("sm_mhz")} '
                  f'{d["clocks"].get("reasons")}  knobs {knobs}')
            if verbose:
                for k, v in d["roofline"]["kernels"].items():
                    rate = f'{v["tflops"]:.0f} TFLOP/s' if "tflops" in v else f'{v["gbps"]:.0f} GB/s'
                    print(f'    {k[:60]:60s} {v["ms"]:.3f} ms  {rate}')


if __name__ == "__main__":
    main()
