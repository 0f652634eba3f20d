import json, sys
d = json.load(open(sys.argv[1]))
print("value %.2f puzzles/s  ms/step %.1f  (%.2f ms per diffusion step)  e2e %.2f  model %.0f TF/s  clocks %s" % (
    d["value"], d["ms_per_step"], d["ms_per_step"] / d["config"]["diffusion_steps"], d["e2e"]["value"], d["model_tflops"], d["clocks"]))
for k, v in d["kernels"].items():
    print("  %-40s %.4f ms  %8.1f %s  frac %.3f" % (k, v["ms"], v["achieved"], v["unit"], v["frac"]))
