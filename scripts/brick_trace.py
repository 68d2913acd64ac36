"""Developer probe: timeline of the brick kernel's last force pass (needs a library built with -DLJ_BRICK_TRACE:
scripts/build_variant.sh trace -DLJ_BRICK_TRACE; LJGPU_LIBRARY=.../libljgpu_trace.so python scripts/brick_trace.py N)."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
lj = importlib.import_module("lennard-jones-live-simulator_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
L = float(np.cbrt(n / 0.8))
p = lj.LennardJonesParameters(cell_length=L, nr_particles=n, kT=1.0, mass=1.0, sigma=1.0, rcut=2.5, epsilon=1.0, tau=0.5, shell=0.4, stepsize=0.001)
lj.host_rng_reset(); st = lj.host_initialize(p)
g = lj.LennardJonesSimulation(p, st, force_kernel=3)
g.integrate_n(1, 400, 0.001)
g.set_profiling(True)                 # plain launches
g.integrate_n(401, 3, 0.001); g.get_energies()
lib = lj.load_library()
path = "/tmp/brick_trace.bin"
rc = lib.ljgpu_debug_trace_dump(path.encode()); assert rc == 0, rc
raw = open(path, "rb").read()
cnt = np.frombuffer(raw[:640], dtype=np.uint32)
ev = np.frombuffer(raw[640:], dtype=np.dtype([("t", "<u8"), ("kind", "<u4"), ("a", "<u4"), ("b", "<u4"), ("warp", "<u4")])).reshape(160, -1)
ctas = [c for c in range(160) if cnt[c]]
t0 = min(ev[c][:cnt[c]]["t"].min() for c in ctas)
starts = np.array([ev[c][:cnt[c]][ev[c][:cnt[c]]["kind"] == 0]["t"][0] - t0 for c in ctas]) / 1e3
ends = np.array([ev[c][:cnt[c]]["t"].max() - t0 for c in ctas]) / 1e3
print(f"N={n}: {len(ctas)} CTAs with events; CTA start spread {starts.min():.1f}..{starts.max():.1f} us, last event at {ends.min():.1f}..{ends.max():.1f} us (mean {ends.mean():.1f})")
nb = [int((ev[c][:cnt[c]]["kind"] == 2).sum()) for c in ctas]
ni = [int((ev[c][:cnt[c]]["kind"] == 3).sum()) for c in ctas]
print("bricks per CTA: min/mean/max", min(nb), np.mean(nb), max(nb), " items per CTA:", min(ni), np.mean(ni), max(ni))
dur = []
for c in ctas:
    e = ev[c][:cnt[c]]
    s3 = {(int(x["a"]), int(x["b"])): int(x["t"]) for x in e[e["kind"] == 3]}
    for x in e[e["kind"] == 4]:
        k = (int(x["a"]), int(x["b"]))
        if k in s3: dur.append((int(x["t"]) - s3[k]) / 1e3)
dur = np.array(dur)
print(f"item duration us: min {dur.min():.2f} median {np.median(dur):.2f} mean {dur.mean():.2f} max {dur.max():.2f}  ({len(dur)} items)")
for c in (ctas[0], ctas[len(ctas) // 2], ctas[int(np.argmax(ends))]):
    e = np.sort(ev[c][:cnt[c]], order="t")
    print(f"--- CTA {c}")
    names = {0: "cta start", 1: "ticket", 2: "tile staged", 3: "item start", 4: "item end", 5: "tables read", 6: "copies issued", 7: "copies landed"}
    for x in (e if len(e) < 80 else e[(e['kind'] != 3) & (e['kind'] != 4)][:40]):
        print(f"   {(int(x['t']) - t0) / 1e3:8.2f} us  warp {int(x['warp']):2d}  {names[int(x['kind'])]:12s} k={int(x['a'])} {'brick' if x['kind'] in (1, 2, 5, 6, 7) else 'chunk'}={int(x['b'])}")
g.close()
