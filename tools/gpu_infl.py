"""Dev script (GPU box): inflation wave + device-resident dynamic cycle timing for a given build of the library.
  python tools/gpu_infl.py <path to .so>"""
import sys, os, json, subprocess
lib = os.path.abspath(sys.argv[1])
code = ("import sys; sys.path.insert(0, '.'); from mesh_navigation_b200 import _lib; _lib.LIB_PATH = %r; sys.argv = ['bench.py', '--batch-goals', '0', '--no-config3', "
        "'--no-cpu-baseline', '--steps', '3']; import runpy; runpy.run_path('bench.py', run_name='__main__')") % lib
r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
try:
    l = json.loads(r.stdout.strip().splitlines()[-1]); o = l["other_kernels"]
    d = o["dynamic_obstacle_update"]
    print(os.path.basename(lib), "inflation", round(o["inflation"]["kernel_ms"], 3), "rounds", o["inflation"]["rounds"], "| dynamic total", round(d["total_wall_ms"], 3),
          "infl_update", round(d["inflation_update_ms"], 3), "kernel", round(d["inflation_kernel_ms"], 3), "layer_changed", round(d["layer_changed_ms"], 3), "equal", d["incremental_equals_full"],
          "| headline ms", round(l["ms_per_step"], 2), "layers", round(o["fused_layers"]["kernel_ms"], 2), flush=True)
except Exception as ex:
    print("failed", ex, r.stdout[-500:], r.stderr[-1500:])
