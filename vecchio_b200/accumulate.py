"""Additive checkpoint / merge of one frame's sample sums.

The pixel is a plain mean over SAMPLES_PER_PIXEL i.i.d. samples (src/main.rs:186-197) and the Philox
counter holds the GLOBAL sample index, so a frame can be rendered as any set of disjoint sample ranges
[begin, begin+count) -- by one GPU over time, by several GPUs, by several runs -- and the partial results
add.  ``vk_render`` with ``spp_begin/spp_count`` returns such a slice already divided by the full ``spp``
(dropped samples stay in the divisor), and its sum-of-squares plane is a plain sum, so both are additive
as they come back.  This module keeps the running total, refuses what would silently double-count
(overlapping ranges, a different frame), and stores it in one ``.npz``.

Nothing here renders: the renderer is whatever object offers ``render(cam, params, want_sumsq)`` the way
``vecchio_b200.Context`` does.
"""
import hashlib
import json
import os

import numpy as np

_FORMAT = 1


def frame_key(scene_name, scene_seed, cam, params, scene_param=0):
    """Identity of a frame: two checkpoints merge only if every field agrees.  ``cam`` is the 24-float
    ``vk_camera`` (hashed), ``params`` a ``vk_render_params`` (its sample range is not part of the identity,
    its variant is not either: the variants produce the same samples)."""
    return {
        "scene": str(scene_name), "scene_seed": int(scene_seed), "scene_param": int(scene_param),
        "camera": hashlib.sha256(bytes(cam)).hexdigest()[:16],
        "width": int(params.width), "height": int(params.height), "spp": int(params.spp),
        "max_depth": int(params.max_depth), "seed": int(params.seed),
        "background": [float(x) for x in params.background],
        # strict and fast arithmetic give statistically equal but different samples: never mix them
        "flags": int(params.flags),
    }


def _merge_ranges(ranges):
    out = []
    for b, e in sorted(ranges):
        if out and b <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], e))
        else:
            out.append((b, e))
    return out


class FrameAccumulator:
    """Running total of one frame: Σ of the partial means (fp64 on the host, so the order of merges does
    not show in the fp32 result), Σ of squares, and the sample ranges they cover."""

    def __init__(self, key, with_sumsq=False):
        self.key = dict(key)
        self.width, self.height, self.spp = key["width"], key["height"], key["spp"]
        shape = (self.height, self.width, 3)
        self.mean_part = np.zeros(shape, dtype=np.float64)
        self.sumsq = np.zeros(shape, dtype=np.float64) if with_sumsq else None
        self.ranges = []  # disjoint, sorted, coalesced [begin, end)
        self.paths = self.rays = self.dropped_samples = 0

    # -- bookkeeping ---------------------------------------------------------------------------------
    @property
    def samples_done(self):
        return sum(e - b for b, e in self.ranges)

    @property
    def complete(self):
        return self.ranges == [(0, self.spp)]

    def missing(self):
        """Sample ranges [begin, end) still to render, in order."""
        out, at = [], 0
        for b, e in self.ranges:
            if b > at:
                out.append((at, b))
            at = e
        if at < self.spp:
            out.append((at, self.spp))
        return out

    def _claim(self, new_ranges):
        for b, e in new_ranges:
            if not (0 <= b < e <= self.spp):
                raise ValueError(f"sample range [{b}, {e}) is outside [0, {self.spp})")
            for b0, e0 in self.ranges:
                if b < e0 and b0 < e:
                    raise ValueError(f"samples [{max(b, b0)}, {min(e, e0)}) are already accumulated: "
                                     "adding them again would count the same Philox samples twice")
        self.ranges = _merge_ranges(self.ranges + list(new_ranges))

    # -- accumulation --------------------------------------------------------------------------------
    def add(self, spp_begin, spp_count, partial_mean, sumsq=None, stats=None):
        """One rendered slice: ``partial_mean`` = what ``render`` returned for samples
        [spp_begin, spp_begin+spp_count) of this frame (Σ_slice / spp)."""
        partial_mean = np.asarray(partial_mean)
        if partial_mean.shape != self.mean_part.shape:
            raise ValueError(f"slice has shape {partial_mean.shape}, the frame {self.mean_part.shape}")
        if (self.sumsq is None) != (sumsq is None):
            raise ValueError("sum-of-squares plane: the accumulator and the slice must both have one or neither")
        self._claim([(int(spp_begin), int(spp_begin) + int(spp_count))])
        self.mean_part += partial_mean
        if sumsq is not None:
            self.sumsq += np.asarray(sumsq)
        if stats is not None:
            self.paths += int(stats.paths)
            self.rays += int(stats.rays)
            self.dropped_samples += int(stats.dropped_samples)

    def merge(self, other):
        """Add another checkpoint of the SAME frame (e.g. from another machine)."""
        if other.key != self.key:
            diff = sorted(k for k in set(self.key) | set(other.key) if self.key.get(k) != other.key.get(k))
            raise ValueError(f"checkpoints are of different frames (differ in {', '.join(diff)})")
        if (self.sumsq is None) != (other.sumsq is None):
            raise ValueError("only one of the checkpoints carries a sum-of-squares plane")
        self._claim(other.ranges)
        self.mean_part += other.mean_part
        if self.sumsq is not None:
            self.sumsq += other.sumsq
        self.paths += other.paths
        self.rays += other.rays
        self.dropped_samples += other.dropped_samples

    # -- results -------------------------------------------------------------------------------------
    def frame(self, allow_partial=False):
        """The frame as ``render`` of all ``spp`` samples returns it: (H, W, 3) float32, row 0 = bottom.  A
        partial total is rescaled to the samples done (a preview) only when asked."""
        if self.complete:
            return self.mean_part.astype(np.float32)
        if not allow_partial:
            raise ValueError(f"frame incomplete: samples {self.missing()} are missing")
        done = self.samples_done
        if done == 0:
            raise ValueError("nothing accumulated yet")
        return (self.mean_part * (self.spp / done)).astype(np.float32)

    def standard_error(self):
        """Per-channel standard error of the pixel mean from Σx and Σx² over the samples done."""
        if self.sumsq is None:
            raise ValueError("no sum-of-squares plane was accumulated")
        n = self.samples_done
        if n < 2:
            raise ValueError("need at least two samples")
        mean = self.mean_part * (self.spp / n)
        var = np.maximum(self.sumsq / n - mean * mean, 0.0) * (n / (n - 1.0))
        return np.sqrt(var / n)

    # -- persistence ---------------------------------------------------------------------------------
    def save(self, path):
        """One ``.npz``, written to a temporary name and renamed so an interrupted save leaves the old file."""
        meta = {"format": _FORMAT, "key": self.key, "ranges": self.ranges, "paths": self.paths, "rays": self.rays,
                "dropped_samples": self.dropped_samples}
        arrays = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), "mean_part": self.mean_part}
        if self.sumsq is not None:
            arrays["sumsq"] = self.sumsq
        tmp = f"{path}.tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            np.savez(f, **arrays)
        os.replace(tmp, path)

    @classmethod
    def load(cls, path):
        with np.load(path) as z:
            meta = json.loads(bytes(z["meta"]).decode())
            if meta.get("format") != _FORMAT:
                raise ValueError(f"{path}: unknown checkpoint format {meta.get('format')}")
            acc = cls(meta["key"], with_sumsq="sumsq" in z.files)
            if z["mean_part"].shape != acc.mean_part.shape:
                raise ValueError(f"{path}: plane shape {z['mean_part'].shape} does not match its header")
            acc.mean_part[...] = z["mean_part"]
            if acc.sumsq is not None:
                acc.sumsq[...] = z["sumsq"]
        acc.ranges = _merge_ranges([(int(b), int(e)) for b, e in meta["ranges"]])
        acc.paths, acc.rays, acc.dropped_samples = int(meta["paths"]), int(meta["rays"]), int(meta["dropped_samples"])
        return acc


def render_resumable(renderer, cam, params, acc, slice_spp, checkpoint_path=None, max_slices=None):
    """Render the samples ``acc`` is missing in slices of at most ``slice_spp``, adding each slice and (if a
    path is given) saving after each.  ``renderer.render(cam, params, want_sumsq)`` is ``Context.render``.
    Returns the number of slices rendered; stops early after ``max_slices``."""
    import copy
    if slice_spp < 1:
        raise ValueError("slice_spp must be positive")
    if (int(params.width), int(params.height), int(params.spp)) != (acc.width, acc.height, acc.spp):
        raise ValueError("render parameters do not match the accumulator's frame")
    n = 0
    for b, e in acc.missing():
        at = b
        while at < e:
            if max_slices is not None and n >= max_slices:
                return n
            count = min(slice_spp, e - at)
            p = copy.copy(params)
            p.spp_begin, p.spp_count = at, count
            rgb, sq, st = renderer.render(cam, p, want_sumsq=acc.sumsq is not None)
            acc.add(at, count, rgb, sq, st)
            if checkpoint_path:
                acc.save(checkpoint_path)
            at += count
            n += 1
    return n
