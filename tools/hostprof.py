import sys, time
sys.path.insert(0, "/root/repo")
import ries_b200
for args in (["3.14159265358979", "-F3"], ["2.5063141592653589", "-l4", "-F3"], ["2.5063141592653589", "-l6", "-F3"]):
    for rep in range(3):
        t0 = time.perf_counter(); ries_b200.call_main(args + ["--no-slow-messages"]); dt = time.perf_counter() - t0
    st = ries_b200.last_stats()
    print(args, "wall %.1f ms  in-main %.1f  engine %.1f  forms %.1f  passes %d  stage_ms %s launches %d" % (
        dt * 1e3, st["seconds"] * 1e3, st["engine_seconds"] * 1e3, st["forms_seconds"] * 1e3, st["passes"],
        [round(x, 2) for x in st["stage_ms"]], st["kernel_launches"]))
