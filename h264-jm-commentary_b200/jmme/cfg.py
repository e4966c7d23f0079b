"""JM `encoder.cfg` files -> jmme_params (SURVEY.md §8(f) rank 4, §5 "config / flags").

JM's lencod reads `Key = value  # comment` lines from a configuration file (`-d encoder.cfg`) and then
applies `-p Key=value` overrides.  Only the keys that bear on motion estimation are mapped; everything else
is carried along untouched in `EncoderCfg.raw`.  Key names are the public JM ones, recalled, not read from
the mounted reference (which has no sources): both the old (`UseHadamard`, `UseFME`, `QPFirstFrame`) and the
newer (`MEDistortionFPel/HPel/QPel`, `Transform8x8Mode`, `ChromaMEEnable`, `SearchMode`, `QPPSlice`) spellings are
understood.
"""
from __future__ import annotations

import re
import shlex
from dataclasses import dataclass, field

from . import abi

_LINE = re.compile(r"^\s*([A-Za-z_][A-Za-z0-9_]*)\s*=\s*(.*?)\s*$")


def parse_encoder_cfg(text: str) -> dict:
    """`Key = value` pairs of a JM configuration text; `#` starts a comment; later keys win; values keep their
    text form (quotes removed).  Lines without `=` are ignored like JM ignores free text."""
    out = {}
    for line in text.splitlines():
        line = line.split("#", 1)[0]
        m = _LINE.match(line)
        if not m:
            continue
        val = m.group(2)
        toks = shlex.split(val) if val else [""]
        out[m.group(1)] = toks[0] if toks else ""
    return out


@dataclass
class EncoderCfg:
    raw: dict = field(default_factory=dict)

    @classmethod
    def load(cls, path, overrides=()):
        """overrides: the `Key=value` strings of lencod's `-p` options."""
        with open(path, "r", errors="replace") as f:
            raw = parse_encoder_cfg(f.read())
        for o in overrides:
            if "=" not in o:
                raise ValueError(f"override {o!r} is not Key=value")
            k, v = o.split("=", 1)
            raw[k.strip()] = v.strip()
        return cls(raw)

    def _int(self, *names, default=None):
        for n in names:
            if n in self.raw:
                try:
                    return int(self.raw[n], 0)
                except ValueError as e:
                    raise ValueError(f"{n} = {self.raw[n]!r} is not an integer") from e
        return default

    # ---- what the ME path needs ------------------------------------------------------------------
    @property
    def input_file(self):
        return self.raw.get("InputFile")

    @property
    def frames(self):
        return self._int("FramesToBeEncoded", default=2)

    def params(self, cost_domain=0) -> dict:
        """Keyword arguments of `Lib.context(...)`.  Raises ValueError for ME modes this library does not
        implement (UMHex / EPZS fast searches, a Hadamard integer-pel stage: SURVEY.md §2.2).
        cost_domain: JM fixes it at compile time (JCOST_CALC_SCALEUP, JM >= 12), so it is an argument, not a key."""
        w, h = self._int("SourceWidth", "OutputWidth"), self._int("SourceHeight", "OutputHeight")
        if not w or not h:
            raise ValueError("SourceWidth / SourceHeight missing")
        mask = 0
        for t, key in enumerate(("InterSearch16x16", "InterSearch16x8", "InterSearch8x16", "InterSearch8x8",
                                 "InterSearch8x4", "InterSearch4x8", "InterSearch4x4"), start=1):
            if self._int(key, default=1):
                mask |= 1 << t
        if not mask:
            raise ValueError("every InterSearch* key is 0: nothing to search")
        # search algorithm: old JM `UseFME` (0 = full), newer `SearchMode` (-1 full, 0 fast full, >0 UMHex / EPZS)
        mode = abi.SEARCH_FASTFULL
        if self._int("UseFME", default=0):
            raise ValueError("UseFME != 0 (UMHexagonS) is not implemented: full search only")
        sm = self._int("SearchMode", default=0)
        if sm > 0:
            raise ValueError(f"SearchMode = {sm} (UMHex / EPZS) is not implemented: -1 (full) or 0 (fast full)")
        if sm < 0 or self._int("UseFastFullSearch", default=1) == 0:
            mode = abi.SEARCH_FULL
        # distortion: `UseHadamard` (old JM: SAD at integer pel, SAD or SATD below) or MEDistortionFPel/HPel/QPel
        # (0 SAD, 1 SSE, 2 Hadamard SAD; JM's own defaults are 0 / 2 / 2)
        had = self._int("UseHadamard", default=None)
        dist = [self._int(k, default=None) for k in ("MEDistortionFPel", "MEDistortionHPel", "MEDistortionQPel")]
        dkw = {}
        if any(d is not None for d in dist):
            dist = [d if d is not None else dflt for d, dflt in zip(dist, (0, 2, 2))]
            if any(d not in (0, 1, 2) for d in dist):
                raise ValueError(f"MEDistortionFPel/HPel/QPel = {dist}: 0 (SAD), 1 (SSE) or 2 (Hadamard SAD)")
            if dist[0] == 2:
                raise ValueError("a Hadamard integer-pel stage (MEDistortionFPel = 2) is not implemented")
            dkw = dict(me_distortion=1, me_distortion_fpel=dist[0], me_distortion_hpel=dist[1], me_distortion_qpel=dist[2])
            had = 1 if dist[1] == 2 else 0
        if self._int("Transform8x8Mode", default=0):
            dkw["transform8x8"] = 1
        ce = self._int("ChromaMEEnable", default=0)
        if ce not in (0, 1):
            raise ValueError(f"ChromaMEEnable = {ce}: 0 or 1 (chroma in the sub-pel stages)")
        if ce:
            dkw["chroma_me"] = 1
        if cost_domain:
            dkw["cost_domain"] = int(cost_domain)
        kw = dict(width=w, height=h, search_range=self._int("SearchRange", default=16),
                  num_refs=min(self._int("NumberReferenceFrames", default=1), abi.MAX_REFS), blocktype_mask=mask,
                  qp=self._int("QPPSlice", "QPRemainingFrame", "QPFirstFrame", default=28),
                  rdopt=1 if self._int("RDOptimization", default=0) else 0,
                  use_hadamard=1 if had in (None, 1) else 0,
                  subpel=0 if self._int("DisableSubpelME", default=0) else 1, search_mode=mode)
        # slices of whole MB rows (SliceMode 1 = fixed number of MBs) -> the in-frame median policy applies per slice
        if self._int("SliceMode", default=0) == 1:
            mbs, mb_w = self._int("SliceArgument", default=0), (w + 15) // 16
            if mbs and mbs % mb_w == 0:
                kw["slice_rows"] = mbs // mb_w
        kw.update(dkw)
        if kw.get("chroma_me") and not kw["subpel"]:
            raise ValueError("ChromaMEEnable = 1 with DisableSubpelME = 1: chroma enters at the sub-pel stages")
        return kw
