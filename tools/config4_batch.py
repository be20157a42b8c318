"""BASELINE configs[4]: a 1050-pair 16 MP batch streamed over every GPU of the box with the C++ driver ugsm_batch
(pair-level data parallel, 4 pairs in flight per GPU), once on the synthetic pair and once from a TIFF image list
through the feeder (ugsm_feeder_*: decode threads -> ring of page-locked pair buffers).  The list holds 4 distinct
16 MP pairs written here as uncompressed RGB TIFFs; like the reference's reader, pairs beyond the list repeat its last
pair.  usage: python tools/config4_batch.py [n_gpus] [pairs]  -> one JSON line per run"""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ug_stereomatcher_b200 import synth
from test_feeder import write_tiff_rgb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1050
batch = os.path.join(ROOT, "ug_stereomatcher_b200", "ugsm_batch")
devs = ",".join(str(i) for i in range(n))
p = subprocess.run([batch, "3264", "4928", str(pairs), "--devices", devs, "--in-flight", "4"], stdout=subprocess.PIPE)
print(p.stdout.decode().strip(), flush=True)
d = tempfile.mkdtemp(prefix="ugsm_tiff_")
names = []
for k in range(4):
    L, R, _, _ = synth.make_pair(4928, 3264, k)
    for tag, img in (("l", L), ("r", R)):
        write_tiff_rgb(os.path.join(d, "%s%d.tif" % (tag, k)), img[:, :, ::-1])
        names.append("%s%d.tif" % (tag, k))
with open(os.path.join(d, "list.xml"), "w") as f:
    f.write('<?xml version="1.0"?>\n<opencv_storage>\n<images>\n%s\n</images>\n</opencv_storage>\n' % "\n".join(names))
for threads in (8, 16, 24):
    p = subprocess.run([batch, "0", "0", str(pairs), "--devices", devs, "--in-flight", "4", "--list",
                        os.path.join(d, "list.xml"), "--decode-threads", str(threads)], stdout=subprocess.PIPE)
    print(p.stdout.decode().strip(), flush=True)
