"""Cold-process time of the `ries` binary (CUDA context creation + search + teardown), what a command-line user sees."""
import subprocess, sys, time
args = sys.argv[1:] or ["2.5063141592653589", "-l6", "-F3"]
for rep in range(4):
    t0 = time.perf_counter()
    r = subprocess.run(["./ries_b200/bin/ries"] + args + ["--no-slow-messages"], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    tline = [l for l in r.stdout.splitlines() if "Time:" in l]
    print(args, "process wall %.3f s" % dt, "| own", tline[0].split("Time:")[1].strip() if tline else "?")
