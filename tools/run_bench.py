"""Tokens per second of the whole token loop (LlamaRunner.run == -[LlamaPredictOperation main], PO.mm:768-901) on the bench
model, with the sampler's candidate stage on the GPU (product) and with the whole sampler on the host (B200_HOST_SAMPLER=1)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import llama_swift_b200 as lsb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-predict", type=int, default=256)
    ap.add_argument("--prompt", default="Once upon a time")
    a = ap.parse_args()
    bench.ensure_model(32)
    path = bench.model_path(32)
    for host_only in ("0", "1", "0", "1"):
        os.environ["B200_HOST_SAMPLER"] = host_only
        stamps = []
        p = lsb.default_run_params(n_predict=a.n_predict, seed=7, n_ctx=512)
        out = lsb.LlamaRunner(path).run(a.prompt, params=p, on_event=lambda kind, piece, code: stamps.append((kind, time.perf_counter())))
        toks = [t for k, t in stamps if k == lsb.EVENT_OUTPUT_TOKEN]
        # steady state: the last 3/4 of the generated tokens
        n0 = len(toks) // 4
        dt = (toks[-1] - toks[n0]) / (len(toks) - 1 - n0)
        g, h = lsb.run_sampler_stats()
        print(f"B200_HOST_SAMPLER={host_only}: {len(out)} tokens, {dt * 1e6:.1f} us/token = {1 / dt:.1f} tok/s  (sampling steps: {g} GPU candidates, {h} host)", flush=True)


if __name__ == "__main__":
    main()
