"""lencod-style front end of the ME path: `python tools/run_yuv.py -d encoder.cfg [-p Key=value ...]`.

Reads the JM configuration file, opens its InputFile (planar YUV 4:2:0) and searches every P frame on the
GPU through the C ABI; prints one line per frame and the sequence throughput."""
import argparse
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
import jmme  # noqa: E402
from jmme import abi  # noqa: E402
from jmme.cfg import EncoderCfg  # noqa: E402
from jmme.sequence import search_sequence, yuv_frames  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("-d", "--cfg", required=True, help="JM encoder.cfg")
ap.add_argument("-p", action="append", default=[], help="Key=value override, as in lencod")
ap.add_argument("--median", action="store_true", help="in-frame median predictors instead of zero predictors")
ap.add_argument("--cost-domain", type=int, default=0, help="1: JM >= 12 scaled-up costs (JCOST_CALC_SCALEUP)")
a = ap.parse_args()
cfg = EncoderCfg.load(a.cfg, a.p)
kw = cfg.params(cost_domain=a.cost_domain)
src = pathlib.Path(cfg.input_file or "")
if not src.is_absolute():
    src = pathlib.Path(a.cfg).resolve().parent / src
lib = jmme.load()
policy = abi.PRED_MEDIAN if a.median else abi.PRED_ZERO
if not a.median:
    kw.pop("slice_rows", None)
w, h = kw.pop("width"), kw.pop("height")
tot_mb, tot_s = 0, 0.0
for n, rec, sec in search_sequence(lib, yuv_frames(src, w, h, cfg.frames, chroma=bool(kw.get("chroma_me"))), policy, **kw):
    c16 = rec["cost"][:, 0].astype(np.int64)
    print(f"frame {n}: {len(rec)} MBs in {1e3 * sec:.2f} ms, mean 16x16 cost {c16.mean():.1f}, "
          f"zero-MV 16x16 blocks {100.0 * np.mean(np.all(rec['mv'][:, 0] == 0, axis=1)):.1f} %")
    if n > 1:                                   # the first searched frame pays the one-time setup
        tot_mb, tot_s = tot_mb + len(rec), tot_s + sec
if tot_s:
    print(f"{tot_mb / tot_s / 1e6:.3f} M MB/s over {tot_mb} MBs (host buffers, reference upload and plane build included)")
