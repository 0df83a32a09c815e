import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ghicp_b200 as g
sc = g.synth.config2()
for rep in range(3):
    reg = g.registration.from_scene(sc, g.FT_BSC, g.CT_NN)
    t0 = time.perf_counter(); reg.build_fd(); dt = (time.perf_counter() - t0) * 1e3
    reg.close()
print("FDTC_DBG", os.environ.get("GHICP_FDTC_DBG"), "POPC", os.environ.get("GHICP_FD_POPC"), "build_fd ms (3rd run)", round(dt, 2))
