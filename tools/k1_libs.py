"""Times K1 (16 MP and 4 MP level sizes) with each tuning build of the library given on the command line."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = ("import sys; sys.path.insert(0, %r)\nfrom ug_stereomatcher_b200 import capi\nh = capi.Handle()\n"
        "print(' '.join('%%dx%%d %%.4f ms' %% (w, hh, h.bench_correlate(w, hh, 10)) for (w, hh) in [(4928, 3264), (2463, 1631), (1231, 815)]))") % root
for lib in sys.argv[1:]:
    env = dict(os.environ, UGSM_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("%-40s %s %s" % (os.path.basename(lib), out.stdout.strip(), out.stderr.strip()[-200:]), flush=True)
