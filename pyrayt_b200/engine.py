"""Device-level driver of the trace kernels: owns the torch buffers, calls the C ABI.

PyTorch is used for device memory, pinned host memory and streams only; every
kernel that runs here is one of ours (libpyrayt_b200.so).  There is no CPU
path: constructing an Engine without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from .scene import FlatScene

_RECORD_MODES = {"all": _lib.RECORD_ALL, "surface": _lib.RECORD_SURFACE, "detector": _lib.RECORD_SURFACE,
                 "none": _lib.RECORD_NONE}


@dataclass
class TraceResult:
    frame: Optional[object]      # torch (15, rows) float64: column-major frame (device, or pinned host)
    rows: int
    counters: dict
    gen_counts: Optional[np.ndarray]  # rows per generation (host int64)
    launches: int                # kernels of ours launched for this trace
    n_leaves: int

    @property
    def ray_surface_tests(self) -> int:
        """sum over generations of live rays x leaf surfaces."""
        return self.counters["generations"] * self.n_leaves


class Engine:
    """One flattened scene resident on one GPU."""

    def __init__(self, scene: FlatScene, device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise _lib.PrtError("pyrayt_b200 needs a CUDA device (B200); there is no CPU fallback")
        self._torch = torch
        self.lib = _lib.load()
        self.scene = scene
        self.device = int(device)
        self.tile = self.lib.prt_tile_rays()
        handle = ctypes.c_void_p()
        desc = scene.as_desc()
        _lib.check(self.lib.prt_scene_create(ctypes.byref(desc), self.device, ctypes.byref(handle)), "prt_scene_create")
        self._handle = handle
        self.n_leaves = scene.n_leaves
        self.rows_per_ray_hint = 4.0
        self._ws = {}
        self._pin = {}
        self.last_transfer = None  # "lean" / "full": how the last host frame crossed the bus
        self._xfer_ms_per_row = {}  # measured once per kind of large host transfer ("lean" / "full")
        self.host_threads = 0  # worker threads of the host-side column rebuild (0 = all hardware threads)

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self.lib.prt_scene_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self):
        return self._torch.device("cuda", self.device)

    def _buf(self, key, numel, dtype):
        """Grow-only workspace tensors (the ABI never allocates)."""
        t = self._ws.get(key)
        if t is not None and t.numel() >= numel and t.dtype == dtype:
            return t
        self._ws[key] = None
        t = None  # drop the old buffer before allocating its replacement
        t = self._torch.empty(int(numel), dtype=dtype, device=self._dev())
        self._ws[key] = t
        return t

    def release_workspace(self) -> None:
        if self._torch.cuda.is_available():
            self._torch.cuda.synchronize(self.device)
        self._ws.clear()
        self._pin.clear()

    # ------------------------------------------------------------------ trace
    def trace(self, d_rays, generation_limit: int = 10, ray_offset: float = 1e-6, record: str = "all",
              detector_sid: int = -1, capacity: Optional[int] = None, to_host: bool = False,
              host_frame=None, zero_copy: bool = False, k1_events=None, method: str = "auto", host_rays=None,
              lean="auto", diagnose: bool = False, precision: str = "fp64") -> TraceResult:
        """Trace a device RaySet.

        precision: "fp64" (the reference's arithmetic, bit-exact against the oracle) or "fp32", the optional fast
        mode (PRT_FLAG_FP32, include/pyrayt_b200.h): the generation loop in single precision, frame values
        within 1e-5 of the scene scale of the FP64 frame; scenes that fit shared memory.

        diagnose: PRT_FLAG_DIAGNOSE -- also count the rays "within 1e-9 of grazing or CSG seams" (counters
        ``grazing_rays`` / ``seam_rays``: the nearest-hit answer of some generation changes when the origin is
        displaced by 1e-9; include/pyrayt_b200.h).  Five searches per generation: a diagnostic, not the fast path.

        method: "single" = one kernel for all generations + ordering pass (staging buffer and frame: 240 B
        of device memory per row); "wavefront" = one launch per generation writing rows in place
        (trace_wavefront: 120 B per row, same frame, a few per cent slower); "auto" = "single" unless
        staging + frame would not fit the device.

        d_rays: torch float64 CUDA tensor (13, N) in the reference RaySet layout.
        to_host: return the frame in pinned host memory (``lean``: "auto" = frames of 2^21 rows or more
        use the lean transfer that rebuilds five columns on the host, see _frame_to_host; True / False force
        it; ``host_rays``: pinned host copy of d_rays if the caller has one; with ``zero_copy`` the gather
        kernel writes straight into the pinned buffer instead).
        k1_events: optional (start, end) torch CUDA events; ``end`` is recorded right after the
        trace kernel so a caller can time that kernel alone on the launching stream.
        """
        torch = self._torch
        assert d_rays.is_cuda and d_rays.dtype == torch.float64 and d_rays.dim() == 2
        assert d_rays.shape[0] == _lib.RAY_ROWS and d_rays.stride(1) == 1
        n = int(d_rays.shape[1])
        stride = int(d_rays.stride(0)) if n > 0 else 0
        G = int(generation_limit)
        mode = _RECORD_MODES[record]
        if method not in ("auto", "single", "wavefront"):
            raise ValueError("method must be 'auto', 'single' or 'wavefront'")
        if precision not in ("fp64", "fp32"):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        fp32 = precision == "fp32"
        flags = (_lib.FLAG_DIAGNOSE if diagnose else 0) | (_lib.FLAG_FP32 if fp32 else 0)
        if (mode != _lib.RECORD_NONE and not zero_copy and k1_events is None and method != "single" and not diagnose
                and not fp32):
            rows_guess = capacity if capacity is not None else min(n * G, n * self.rows_per_ray_hint * 1.05)
            total = torch.cuda.get_device_properties(self.device).total_memory
            if method == "wavefront" or 8 * (_lib.STAGE_COLS + _lib.FRAME_COLS) * rows_guess > 0.8 * total:
                return self.trace_wavefront(d_rays, generation_limit=G, ray_offset=ray_offset, record=record,
                                            detector_sid=detector_sid, capacity=capacity, to_host=to_host,
                                            host_frame=host_frame, host_rays=host_rays, lean=lean)
        n_tiles = max(1, (n + self.tile - 1) // self.tile)
        launches = 0
        with torch.cuda.device(self.device):
            ctr = self._buf("ctr", _lib.COUNTER_WORDS, torch.int64)
            params = _lib.PrtParams(G, mode, flags, 0, float(ray_offset), int(detector_sid))
            if mode == _lib.RECORD_NONE:
                ctr.zero_()
                _lib.check(self.lib.prt_trace(self._handle, ctypes.byref(params), d_rays.data_ptr(), n, stride,
                                              None, ctr.data_ptr(), self._stream()), "prt_trace")
                launches += 1
                counters = self._counters(ctr)
                return TraceResult(None, 0, counters, None, launches, self.n_leaves)

            if G * n_tiles > (1 << 31):
                raise _lib.PrtError("generation_limit x tiles too large for one call; trace the rays in chunks")
            cap = int(capacity) if capacity is not None else int(min(n * G, max(n * self.rows_per_ray_hint * 1.05, 4096)))
            cap = max(cap, 1)
            late_gather = to_host and zero_copy  # the gather writes into a host frame sized from the row count
            if late_gather and fp32:
                raise _lib.PrtError("zero_copy is an FP64-path experiment; use to_host=True with precision='fp32'")
            while True:
                stage = self._buf("stage", _lib.STAGE_COLS * cap, torch.float64)
                run_start = self._buf("run_start", G * n_tiles, torch.int64)
                run_count = self._buf("run_count", G * n_tiles, torch.int32)
                run_base = self._buf("run_base", G * n_tiles, torch.int64)
                gen_off = self._buf("gen_off", G + 1, torch.int64)
                ctr.zero_()
                run_count[: G * n_tiles].zero_()
                rec = _lib.PrtRecords(stage.data_ptr(), cap, run_start.data_ptr(), run_count.data_ptr(),
                                      run_base.data_ptr(), n_tiles)
                _lib.check(self.lib.prt_trace(self._handle, ctypes.byref(params), d_rays.data_ptr(), n, stride,
                                              ctypes.byref(rec), ctr.data_ptr(), self._stream()), "prt_trace")
                launches += 1
                if k1_events is not None:
                    k1_events[1].record()
                _lib.check(self.lib.prt_scan_runs(ctypes.byref(rec), G, gen_off.data_ptr(), self._stream()),
                           "prt_scan_runs")
                launches += 2
                frame = None
                if not late_gather:
                    # the ordering pass is enqueued before anything is read back: the frame is sized like the
                    # staging buffer (rows <= cap whenever nothing was dropped), the kernel takes the row
                    # offsets from device memory, and the step has one host synchronisation, at its end
                    frame = torch.empty((_lib.FRAME_COLS, cap), dtype=torch.float64, device=self._dev())
                    self._gather_into(rec, d_rays, G, gen_off, frame, cap,
                                      layout=_lib.LAYOUT_FP32_RECORDS if fp32 else 0)
                    launches += 1
                host = torch.cat([ctr, gen_off[: G + 1]]).cpu()  # one small D2H + sync
                counters = dict(zip(_lib.COUNTER_FIELDS, (int(x) for x in host[: len(_lib.COUNTER_FIELDS)])))
                goff = host[_lib.COUNTER_WORDS:].numpy()
                if counters["rows_dropped"] == 0:
                    break
                # staging overflowed: the kernel kept counting; retry once with the exact size
                frame = None
                cap = int(counters["rows_reserved"])
            rows = int(goff[G])
            if n > 0:
                self.rows_per_ray_hint = max(self.rows_per_ray_hint, rows / n)
            gen_counts = np.diff(goff).astype(np.int64)
            if late_gather:
                frame = self._gather(rec, G, gen_off, rows, to_host, host_frame, zero_copy, goff, d_rays, host_rays, lean)
                launches += 1 if rows else 0
            elif rows == 0:
                frame = torch.empty((_lib.FRAME_COLS, 0), dtype=torch.float64, device=None if to_host else self._dev())
            else:
                frame = frame[:, :rows]
                if to_host:
                    frame = self._frame_to_host(frame, rows, goff, d_rays, host_frame, host_rays, lean)
            return TraceResult(frame, rows, counters, gen_counts, launches, self.n_leaves)

    def _gather_into(self, rec, d_rays, G, gen_off, frame, frame_capacity, layout: int = 0):
        """prt_gather_frame: expand the staged records of `rec` into `frame` (device or pinned host)."""
        n = int(d_rays.shape[1])
        _lib.check(self.lib.prt_gather_frame(self._handle, ctypes.byref(rec), d_rays.data_ptr(), n,
                                             int(d_rays.stride(0)) if n else 0, G, gen_off.data_ptr(),
                                             frame.data_ptr(), int(frame.stride(0)), int(frame_capacity), layout,
                                             self._stream()), "prt_gather_frame")

    # ------------------------------------------------------------------ large ray sets: one generation per launch
    WAVEFRONT_MIN_RAYS = 1 << 18

    def trace_wavefront(self, d_rays, generation_limit: int = 10, ray_offset: float = 1e-6, record: str = "all",
                        detector_sid: int = -1, capacity: Optional[int] = None, to_host: bool = False,
                        host_frame=None, nearest_events=None, host_rays=None, lean="auto") -> TraceResult:
        """Same result as trace(), computed generation by generation (prt_trace_wavefront): every row is
        written straight to its final frame position, so there is no staging buffer and no ordering
        pass.  The device frame is a (15, rows) view of a (15, capacity) buffer.

        nearest_events: optional list that receives (start, end) torch CUDA events around each launch
        of the step kernel (generation_limit + 1 of them; bench.py times the dominant kernel with them).
        """
        torch = self._torch
        assert d_rays.is_cuda and d_rays.dtype == torch.float64 and d_rays.dim() == 2
        assert d_rays.shape[0] == _lib.RAY_ROWS and d_rays.stride(1) == 1
        n = int(d_rays.shape[1])
        stride = int(d_rays.stride(0)) if n > 0 else 0
        G = int(generation_limit)
        mode = _RECORD_MODES[record]
        if mode == _lib.RECORD_NONE:
            raise _lib.PrtError("trace_wavefront records rows; use trace(record='none') for counters only")
        wtile = self.lib.prt_wave_tile()
        n_tiles = max(1, (n + wtile - 1) // wtile)
        with torch.cuda.device(self.device):
            ctr = self._buf("ctr", _lib.COUNTER_WORDS, torch.int64)
            gen_off = self._buf("gen_off", G + 1, torch.int64)
            ws = _lib.PrtWaveWorkspace(
                self._buf("w_state", 7 * max(n, 1), torch.float64).data_ptr(),
                self._buf("w_flag", max(n, 1), torch.int32).data_ptr(),
                self._buf("w_hit_t", max(n, 1), torch.float64).data_ptr(),
                self._buf("w_hit_leaf", max(n, 1), torch.int32).data_ptr(),
                self._buf("w_tile_count", 2 * n_tiles, torch.int32).data_ptr(),
                self._buf("w_tile_base", 2 * n_tiles, torch.int64).data_ptr(),
                self._buf("w_alive", G + 1, torch.int64).data_ptr(), n_tiles)
            params = _lib.PrtParams(G, mode, 0, 0, float(ray_offset), int(detector_sid))
            events = None
            if nearest_events is not None:
                pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                         for _ in range(G + 1)]
                for e0, e1 in pairs:  # a torch event gets its CUDA handle at its first record()
                    e0.record()
                    e1.record()
                events = (ctypes.c_void_p * (2 * G + 2))(*[ev.cuda_event for pair in pairs for ev in pair])
                nearest_events[:] = pairs
            cap = int(capacity) if capacity is not None else int(min(n * G, max(n * self.rows_per_ray_hint * 1.02, 4096)))
            cap = max(cap, 1)
            frame = None
            while True:
                frame = None  # (a retry frees the short buffer first)
                frame = torch.empty((_lib.FRAME_COLS, cap), dtype=torch.float64, device=self._dev())
                ctr.zero_()
                _lib.check(self.lib.prt_trace_wavefront(self._handle, ctypes.byref(params), d_rays.data_ptr(), n, stride,
                                                        ctypes.byref(ws), frame.data_ptr(), cap, cap,
                                                        gen_off.data_ptr(), ctr.data_ptr(), events, self._stream()),
                           "prt_trace_wavefront")
                host = torch.cat([ctr, gen_off[: G + 1]]).cpu()  # one small D2H + sync
                counters = dict(zip(_lib.COUNTER_FIELDS, (int(x) for x in host[: len(_lib.COUNTER_FIELDS)])))
                goff = host[_lib.COUNTER_WORDS:].numpy()
                if counters["rows_dropped"] == 0:
                    break
                cap = int(counters["rows_reserved"])  # the kernels kept counting: retry once with the exact size
            rows = int(goff[G])
            if n > 0:
                self.rows_per_ray_hint = max(self.rows_per_ray_hint, rows / n)
            gen_counts = np.diff(goff).astype(np.int64)
            launches = 2 * G + 1
            out = frame[:, :rows]
            if to_host:
                out = self._frame_to_host(frame, rows, goff, d_rays, host_frame, host_rays, lean) if rows else \
                    torch.empty((_lib.FRAME_COLS, 0), dtype=torch.float64)
            return TraceResult(out, rows, counters, gen_counts, launches, self.n_leaves)

    # ------------------------------------------------------------------ many small traces (N4)
    SMALL_MAX_RAYS = 4096
    SMALL_MAX_ROWS = 1 << 16

    def small_fits(self, n: int, generation_limit: int) -> bool:
        return 0 < n <= self.SMALL_MAX_RAYS and n * int(generation_limit) <= self.SMALL_MAX_ROWS

    def small_ray_buffer(self, n: int):
        """Persistent (13, n) device RaySet for trace_small: fill it (sources, or one H2D copy) and call
        trace_small(n, ...).  Its address is part of the captured launch sequence."""
        torch = self._torch
        t = self._ws.get(("small_rays", n))
        if t is None:
            t = torch.zeros((_lib.RAY_ROWS, n), dtype=torch.float64, device=self._dev())
            self._ws[("small_rays", n)] = t
        return t

    def trace_small(self, n: int, generation_limit: int = 10, ray_offset: float = 1e-6, record: str = "all",
                    detector_sid: int = -1, use_graph: bool = True) -> TraceResult:
        """Latency path for the optimiser loops of examples/lens_design.ipynb (cells 28-33: thousands of
        traces of <= 21 rays through a system whose radii change in between).

        The staging buffer is sized for the worst case (n x generation_limit rows), so nothing depends
        on a count read back mid-way: clear -> trace -> scan -> gather -> two D2H copies are enqueued
        back to back, captured once into a CUDA graph per (n, generation_limit, record mode) and
        replayed with one launch and one synchronisation per trace.  The scene is read from its device
        blob at run time, so prt_scene_update between replays is seen by the next one.
        """
        torch = self._torch
        G = int(generation_limit)
        if not self.small_fits(n, G):
            raise _lib.PrtError("trace_small is for small ray sets; use trace()")
        mode = _RECORD_MODES[record]
        if mode == _lib.RECORD_NONE:
            raise _lib.PrtError("trace_small records rows; use trace(record='none') for counters only")
        key = ("small", n, G, mode, int(detector_sid), float(ray_offset))
        st = self._ws.get(key)
        with torch.cuda.device(self.device):
            if st is None:
                st = self._small_state(n, G, mode, detector_sid, ray_offset)
                self._ws[key] = st
            if use_graph and st["graph"] is None and st["uses"] >= 1:
                # capture on the second use: the first one ran eagerly (module load, attribute calls)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._small_enqueue(st)
                st["graph"] = g
            if use_graph and st["graph"] is not None:
                st["graph"].replay()
            else:
                self._small_enqueue(st)
            st["uses"] += 1
            torch.cuda.current_stream(self.device).synchronize()
        meta = st["h_meta"]
        counters = dict(zip(_lib.COUNTER_FIELDS, (int(x) for x in meta[: len(_lib.COUNTER_FIELDS)])))
        goff = meta[_lib.COUNTER_WORDS: _lib.COUNTER_WORDS + G + 1].numpy()
        rows = int(goff[G])
        frame = st["h_frame"][:, :rows].clone()  # the pinned buffer is overwritten by the next replay
        return TraceResult(frame, rows, counters, np.diff(goff).astype(np.int64), st["launches"], self.n_leaves)

    def _small_state(self, n, G, mode, detector_sid, ray_offset):
        torch = self._torch
        dev = self._dev()
        cap = n * G
        n_tiles = max(1, (n + self.tile - 1) // self.tile)
        # [counters | generation offsets] and the run counts are cleared / read back together
        meta = torch.zeros(_lib.COUNTER_WORDS + G + 1, dtype=torch.int64, device=dev)
        run_count = torch.zeros(G * n_tiles, dtype=torch.int32, device=dev)
        st = {
            "n": n, "G": G, "cap": cap, "meta": meta, "run_count": run_count,
            "run_start": torch.zeros(G * n_tiles, dtype=torch.int64, device=dev),
            "run_base": torch.zeros(G * n_tiles, dtype=torch.int64, device=dev),
            "stage": torch.empty(_lib.STAGE_COLS * cap, dtype=torch.float64, device=dev),
            "frame": torch.zeros((_lib.FRAME_COLS, cap), dtype=torch.float64, device=dev),
            "h_meta": torch.zeros(_lib.COUNTER_WORDS + G + 1, dtype=torch.int64).pin_memory(),
            "h_frame": torch.zeros((_lib.FRAME_COLS, cap), dtype=torch.float64).pin_memory(),
            "rays": self.small_ray_buffer(n),
            "params": _lib.PrtParams(G, mode, 0, 0, float(ray_offset), int(detector_sid)),
            "graph": None, "uses": 0, "launches": 4,
        }
        st["rec"] = _lib.PrtRecords(st["stage"].data_ptr(), cap, st["run_start"].data_ptr(), run_count.data_ptr(),
                                    st["run_base"].data_ptr(), n_tiles)
        return st

    def _small_enqueue(self, st):
        n, G = st["n"], st["G"]
        meta, rays = st["meta"], st["rays"]
        stream = self._stream()
        meta.zero_()
        st["run_count"].zero_()
        _lib.check(self.lib.prt_trace(self._handle, ctypes.byref(st["params"]), rays.data_ptr(), n, int(rays.stride(0)),
                                      ctypes.byref(st["rec"]), meta.data_ptr(), stream), "prt_trace")
        gen_off = meta[_lib.COUNTER_WORDS:]
        _lib.check(self.lib.prt_scan_runs(ctypes.byref(st["rec"]), G, gen_off.data_ptr(), stream), "prt_scan_runs")
        self._gather_into(st["rec"], rays, G, gen_off, st["frame"], st["cap"])
        st["h_meta"].copy_(meta, non_blocking=True)
        st["h_frame"].copy_(st["frame"], non_blocking=True)

    # ------------------------------------------------------------------ device frame -> pinned host frame
    LEAN_MIN_ROWS = 1 << 21

    def _pinned(self, key, numel, dtype):
        """Grow-only pinned host buffers (kept apart from the device workspace)."""
        t = self._pin.get(key)
        if t is not None and t.numel() >= numel and t.dtype == dtype:
            return t
        self._pin[key] = None
        t = self._torch.empty(int(numel), dtype=dtype, pin_memory=True)
        self._pin[key] = t
        return t

    def _frame_to_host(self, frame, rows, goff, d_rays, host_frame=None, host_rays=None, lean="auto"):
        """Copy a device frame (15, >= rows; contiguous columns) into pinned host memory.

        Large frames can take the lean transfer (csrc/prt_transfer.cu): five columns are rebuilt on the
        host from the input rays while the other ten stream over the bus.  Whether that beats copying all
        fifteen depends on the host (cores per GPU, memory bandwidth left beside the DMA traffic), so with
        lean="auto" an engine times its first large transfer of each kind and keeps the faster one.
        host_rays: pinned host copy of the (13, n) RaySet if the caller has one (else the four rows needed
        are copied back).
        """
        torch = self._torch
        out = host_frame if host_frame is not None else torch.empty(
            (_lib.FRAME_COLS, rows), dtype=torch.float64, pin_memory=True)
        assert out.shape[0] == _lib.FRAME_COLS and out.shape[1] >= rows and out.stride(1) == 1
        stream = torch.cuda.current_stream(self.device)
        n = int(d_rays.shape[1]) if d_rays is not None else 0
        large = n > 0 and rows >= self.LEAN_MIN_ROWS
        calibrating = False
        if lean == "auto" and large:
            untried = [m for m in ("lean", "full") if m not in self._xfer_ms_per_row]
            calibrating = bool(untried)
            mode = untried[0] if untried else min(self._xfer_ms_per_row, key=self._xfer_ms_per_row.get)
        else:
            mode = "lean" if (lean is True and n > 0) else "full"
        self.last_transfer = mode
        if mode == "lean":  # buffers first, so that a calibration run does not time their allocation
            packed = torch.empty(rows, dtype=torch.int64, device=self._dev())
            bad = self._buf("pack_bad", 1, torch.int64)
            h_packed = self._pinned("h_packed", rows, torch.int64)
            h_bad = self._pinned("h_bad", 1, torch.int64)
            if host_rays is None:
                h4 = self._pinned("h_ray_rows", 4 * n, torch.float64).view(-1)[: 4 * n].view(4, n)
        if calibrating:
            stream.synchronize()
            t_begin = time.perf_counter()
        if mode == "full":
            for c in range(_lib.FRAME_COLS):  # column by column: contiguous on both sides whatever the strides
                out[c, :rows].copy_(frame[c, :rows], non_blocking=True)
            stream.synchronize()
        else:
            d_goff = self._buf("pack_goff", len(goff), torch.int64)
            d_goff[: len(goff)].copy_(torch.from_numpy(np.ascontiguousarray(goff, dtype=np.int64)), non_blocking=True)
            _lib.check(self.lib.prt_frame_pack(frame.data_ptr(), rows, int(frame.stride(0)), d_rays.data_ptr(), n,
                                               int(d_rays.stride(0)), d_goff.data_ptr(), len(goff) - 1,
                                               packed.data_ptr(), bad.data_ptr(), self._stream()), "prt_frame_pack")
            h_packed[:rows].copy_(packed, non_blocking=True)
            h_bad.copy_(bad, non_blocking=True)
            if host_rays is None:
                for k, row in enumerate((8, 9, 10, 12)):
                    h4[k].copy_(d_rays[row], non_blocking=True)
                ray_rows = [h4[k] for k in range(4)]
            else:
                ray_rows = [host_rays[row] for row in (8, 9, 10, 12)]
                assert all(r.stride(0) == 1 and r.dtype == torch.float64 for r in ray_rows)
            arrived = torch.cuda.Event()
            arrived.record(stream)
            for c in (3, 6, 7, 8, 9, 10, 11, 12, 13, 14):  # the columns only the device knows
                out[c, :rows].copy_(frame[c, :rows], non_blocking=True)
            arrived.synchronize()
            if int(h_bad[0]) == 0:
                goff = np.ascontiguousarray(goff, dtype=np.int64)
                _lib.check(self.lib.prt_host_expand_frame(
                    h_packed.data_ptr(), rows, goff.ctypes.data_as(ctypes.c_void_p), len(goff) - 1,
                    ray_rows[0].data_ptr(), ray_rows[1].data_ptr(), ray_rows[2].data_ptr(), ray_rows[3].data_ptr(),
                    out.data_ptr(), int(out.stride(0)), int(self.host_threads)), "prt_host_expand_frame")
            else:  # some row does not verify (ids not consecutive, exotic surface ids): copy the five columns too
                for c in (0, 1, 2, 4, 5):
                    out[c, :rows].copy_(frame[c, :rows], non_blocking=True)
                self.last_transfer = "full"
            stream.synchronize()
        if calibrating:
            self._xfer_ms_per_row[mode] = 1e3 * (time.perf_counter() - t_begin) / rows
        return out[:, :rows]

    def _gather(self, rec, G, gen_off, rows, to_host, host_frame, zero_copy, goff=None, d_rays=None, host_rays=None,
                lean="auto"):
        torch = self._torch
        if rows == 0:
            if to_host:
                return torch.empty((_lib.FRAME_COLS, 0), dtype=torch.float64)
            return torch.empty((_lib.FRAME_COLS, 0), dtype=torch.float64, device=self._dev())
        if to_host and zero_copy:
            out = host_frame if host_frame is not None else torch.empty(
                (_lib.FRAME_COLS, rows), dtype=torch.float64, pin_memory=True)
            assert out.is_pinned() and out.shape[0] == _lib.FRAME_COLS and out.shape[1] >= rows
            self._gather_into(rec, d_rays, G, gen_off, out, rows)
            torch.cuda.current_stream(self.device).synchronize()
            return out[:, :rows]
        frame = torch.empty((_lib.FRAME_COLS, rows), dtype=torch.float64, device=self._dev())
        self._gather_into(rec, d_rays, G, gen_off, frame, rows)
        if not to_host:
            return frame
        return self._frame_to_host(frame, rows, goff, d_rays, host_frame, host_rays, lean)

    def _counters(self, ctr) -> dict:
        host = ctr.cpu()
        return dict(zip(_lib.COUNTER_FIELDS, (int(x) for x in host[: len(_lib.COUNTER_FIELDS)])))

    # ------------------------------------------------------------------ scene updates / nearest hit
    def update_scene(self, scene: FlatScene) -> None:
        """Re-encode a (moved / re-parameterised) scene into this engine without re-creating it."""
        desc = scene.as_desc()
        same_structure = (np.array_equal(scene.node_kind, self.scene.node_kind)
                          and np.array_equal(scene.comp_node_begin, self.scene.comp_node_begin)
                          and np.array_equal(scene.leaf_type, self.scene.leaf_type))
        if not same_structure:  # captured launch sequences hold the blob size and the kernel variant
            for st in [v for k, v in self._ws.items() if isinstance(k, tuple) and k[0] == "small"]:
                st["graph"], st["uses"] = None, 0
        with self._torch.cuda.device(self.device):
            _lib.check(self.lib.prt_scene_update(self._handle, ctypes.byref(desc), self._stream()), "prt_scene_update")
        self.scene = scene
        self.n_leaves = scene.n_leaves

    def nearest_hit(self, d_rays, normals: bool = False, renderer: bool = False):
        """_st_propagate alone: (distance (N,), surface id (N,), world normals (3,N) or None).

        renderer=True: the variant the reference's renderers run (prt_render_hit): a ray whose hits on a
        component are all behind it picks up that component's first, negative, hit."""
        torch = self._torch
        entry, what = (self.lib.prt_render_hit, "prt_render_hit") if renderer else \
            (self.lib.prt_nearest_hit, "prt_nearest_hit")
        r = d_rays.reshape(8, -1).contiguous()
        n = int(r.shape[1])
        t = torch.empty(n, dtype=torch.float64, device=self._dev())
        sid = torch.empty(n, dtype=torch.int64, device=self._dev())
        nrm = torch.empty((3, n), dtype=torch.float64, device=self._dev()) if normals else None
        with torch.cuda.device(self.device):
            _lib.check(entry(self._handle, r.data_ptr(), n, t.data_ptr(), sid.data_ptr(),
                             nrm.data_ptr() if normals else None, self._stream()), what)
        return t, sid, nrm

    # ------------------------------------------------------------------ component.intersect
    def intersect(self, component: int, d_rays):
        """component.intersect(rays (2,4,N) device) -> (hits (m,N), surface ids (m,N)) device tensors."""
        torch = self._torch
        r = d_rays.reshape(8, -1).contiguous()
        n = int(r.shape[1])
        m = self.scene.component_slots(component)
        hits = torch.empty((m, n), dtype=torch.float64, device=self._dev())
        sids = torch.empty((m, n), dtype=torch.int64, device=self._dev())
        slots = ctypes.c_int32()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.prt_intersect(self._handle, component, r.data_ptr(), n, hits.data_ptr(),
                                              sids.data_ptr(), ctypes.byref(slots), self._stream()), "prt_intersect")
        assert slots.value == m
        return hits, sids
