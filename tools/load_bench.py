"""Load-path timing: llama_model_load of the 7B bench file -- memory-mapped part files, pinned
double-buffered staging, GPU-side repack into the decode stream and the prefill tile layout.  Prints seconds for a load
with the file in the page cache (it was just written / read) and, when permitted, after dropping the page cache."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import llama_swift_b200 as lsb

path = bench.ensure_model(32)
size = os.path.getsize(path)
for label in ("page cache warm", "page cache warm (2nd)", "after drop_caches"):
    if label == "after drop_caches":
        try:
            os.sync()
            open("/proc/sys/vm/drop_caches", "w").write("3\n")
        except OSError as e:
            print(f"drop_caches not permitted ({e}); skipping the cold-file load")
            break
    t0 = time.perf_counter()
    m = lsb.llama_model_load(path, n_ctx=520)
    dt = time.perf_counter() - t0
    print(f"load 7B Q4_0 ({size / 1e9:.2f} GB file), {label}: {dt:.2f} s  ({size / dt / 1e9:.2f} GB/s incl. both GPU layouts)", flush=True)
    m.free()
for env, label in ((("B200_PREFILL_COPY", "0"),), "without the prefill (tcgen05) weight copy"),:
    for k, v in env: os.environ[k] = v
    t0 = time.perf_counter()
    m = lsb.llama_model_load(path, n_ctx=520)
    dt = time.perf_counter() - t0
    print(f"load 7B Q4_0, page cache warm, {label}: {dt:.2f} s", flush=True)
    m.free()
t0 = time.perf_counter(); a = lsb.llama_model_acquire(path, n_ctx=520); t1 = time.perf_counter(); a.release()
t2 = time.perf_counter(); b = lsb.llama_model_acquire(path, n_ctx=520); t3 = time.perf_counter(); b.release(); lsb.llama_model_cache_clear()
print(f"acquire (first) {t1 - t0:.2f} s, re-acquire of the resident model {(t3 - t2) * 1e3:.3f} ms")
