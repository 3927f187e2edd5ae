#!/usr/bin/env python
"""Benchmark of the B200 interpolation path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Primary line (BASELINE.json metric, second half): intertable APPLY in algorithmic GB/s against the
HBM roofline.  Workload = configs[3] of BASELINE.json as one GPU of the 8xB200 box sees it: the
51-member x 60-step ENS batch (3060 messages) sharded by member -> 383 O1280 messages per GPU,
applied to the EFAS 5 km grid through an invdist (k=4) table, with missing values propagated: the
resident fields carry the marker NaN where the GRIB bitmap says "missing" (folded in by the kernel
that produces them: the PCIe gather / g2p_manip / g2p_mark_masked), the apply kernel returns the
fields (masked cells = mv_out) and a bit-packed mask.  One "step" = one pass over the rank's resident
batch = ONE launch of the apply kernel.  Further objects of the same JSON line:
  `build`     configs[2]: intertable BUILD target-pts/s, O1280 -> GloFAS 0.05 deg (21.6 M targets, invdist k=4),
              targets sharded over ranks
  `configs0`  configs[0]: O320 -> EFAS invdist k=4, GPU build + single-message apply beside the CPU reference
  `configs1`  configs[1]: O1280 -> GloFAS 0.1 deg nearest, build + single-message apply with its own roofline
  `pipeline`  configs[4]: x*1000 -> 24 h accumulation -> cut-off -> apply -> lapse-rate correction, device-resident
              and from host buffers with masks (its own e2e)
  `adw`       SURVEY 8(f)-4: mode adw (Shepard, k=11) build

`--impl reference` times the reference's CPU algorithm (oracle/ restatement: np.append + einsum per
message; cKDTree(leafsize=30, workers=-1) + numpy weights for the build) on bounded samples of the
same workloads on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from pyg2p_b200 import synth  # noqa: E402

K_INVDIST = 4
MESSAGES_FULL_JOB = 51 * 60


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def algorithmic_bytes_apply(b, n_touched, m, k, masks):
    """8*B*(N_touched + M) + M*k*(4+8) [+ B*ceil(N_touched/8) mask bits]."""
    total = 8 * b * (n_touched + m) + m * k * (4 + (8 if k > 1 else 0))
    if masks:
        total += b * ((n_touched + 7) // 8)
    return total


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  The timed region
    of the apply step is ~20 ms, shorter than one `nvidia-smi -lms` period, so the samples come from NVML directly
    (the library nvidia-smi itself reads), polled by a thread every ~2 ms between `start()` and `stop()`; `nvidia-smi`
    is the fall-back when NVML cannot be loaded."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []          # nvidia-smi fall-back: csv rows
        self.samples = []       # NVML: (sm_mhz, power_w, reasons bit mask)
        self.proc = None
        self.nvml = None
        self.handle = None
        self.sm_max = None
        self._stop = threading.Event()
        self.thread = None

    def _open_nvml(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            handle = None
            try:    # the CUDA device behind `gpu_index` (CUDA_VISIBLE_DEVICES may renumber): look it up by UUID
                uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith('GPU-') else 'GPU-' + uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self.nvml, self.handle = pynvml, handle
            return True
        except Exception:
            self.nvml = None
            return False

    def _poll(self):
        nv, h = self.nvml, self.handle
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = float('nan')
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def prepare(self):
        """load NVML ahead of the timed region (nvmlInit takes tens of ms)"""
        self._open_nvml()

    def start(self):
        if self.nvml is not None or self._open_nvml():
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            nv = self.nvml
            bits = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
                    'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
            sm = [s[0] for s in self.samples]
            pw = [s[1] for s in self.samples if s[1] == s[1]]
            reasons = sorted(nm for nm, bit in bits.items() if any(s[2] & bit for s in self.samples))
            return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.sm_max,
                    'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': reasons,
                    'source': 'NVML polled every ~2 ms over the timed region'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons),
                'source': 'nvidia-smi -lms 100'}


# =============================================================================== reference arm (CPU)
_O1280_ORACLE = {}


def _o1280_oracle(mode, k):
    """the reference's `ScipyInterpolation.__init__` on the O1280 source (cKDTree + self query), once per process;
    returns (oracle, seconds of the fixed part)."""
    from oracle import interp_ref as R
    key = (mode, k)
    if key not in _O1280_ORACLE:
        lat, lon, det = synth.octahedral_grid(1280)
        t0 = time.perf_counter()
        o = R.BuildOracle(lon, lat, det, np.zeros(lat.size), k, synth.NC_FILL_F64, synth.GRIB_MV, parallel=True, mode=mode)
        _O1280_ORACLE[key] = (o, time.perf_counter() - t0)
    return _O1280_ORACLE[key]


def cpu_apply_reference(n_messages, seed=0):
    """The reference's apply (`_interpolate_scipy_append_mv`, one message per call, single thread) on a
    bounded sample of the workload.  Returns (seconds, algorithmic bytes, sample description)."""
    import numpy.ma as ma

    from oracle import interp_ref as R
    lat, lon, det = synth.octahedral_grid(1280)
    tl, tlo = synth.efas_target()
    t0 = time.perf_counter()
    o, _ = _o1280_oracle('invdist', K_INVDIST)
    _, w, idx = o.interpolate(tlo, tl)
    t_table = time.perf_counter() - t0
    rec = R.to_record(idx, w)                      # strided views, as `_read_intertable` hands them over
    n_touched = int(np.unique(idx[idx < lat.size]).size)
    mask = synth.missing_mask(lat)
    fields = [ma.masked_where(mask, np.where(mask, synth.GRIB_MV, synth.field_2t(lat, lon, member=i)), copy=False)
              for i in range(min(4, n_messages))]
    t0 = time.perf_counter()
    for i in range(n_messages):
        with np.errstate(all='ignore'):
            R.apply_as_reference(fields[i % len(fields)], rec['coeffs'], rec['indexes'], K_INVDIST, synth.NC_FILL_F64)
    dt = time.perf_counter() - t0
    nbytes = algorithmic_bytes_apply(n_messages, n_touched, tl.size, K_INVDIST, True)
    return dt, nbytes, {'messages': n_messages, 'n_touched': n_touched, 'table_build_s': round(t_table, 3),
                        'what': f'{n_messages} of the 383 O1280 messages, np.append + einsum per message, 1 thread'}


def cpu_build_reference(row_fraction=8):
    """The reference's build (`ScipyInterpolation`, -X => cKDTree workers=-1) for O1280 -> 0.05 deg invdist on
    1/row_fraction of the target rows.  Fixed costs (tree, self query) are paid in full."""
    tl, tlo = synth.latlon_target(0.05)
    rows = tl.shape[0] // row_fraction
    tl, tlo = tl[:rows], tlo[:rows]
    o, t_fixed = _o1280_oracle('invdist', K_INVDIST)
    t0 = time.perf_counter()
    o.interpolate(tlo, tl)
    t_sample = time.perf_counter() - t0
    m_full = 7200 * 3000
    est_full = t_fixed + t_sample * (m_full / tl.size)
    return {'metric': 'intertable build target-pts/s', 'value': m_full / est_full, 'unit': 'target-pts/s',
            'kind': 'port', 'cores': os.cpu_count(),
            'sample': f'{tl.size} of {m_full} targets (first {rows} rows); tree + self-query {t_fixed:.2f} s paid in '
                      f'full, target part {t_sample:.2f} s scaled x{row_fraction}; cKDTree workers=-1, numpy weights',
            'seconds_full_estimate': est_full, 'seconds_fixed': t_fixed, 'seconds_sample': t_sample}


def cpu_configs0():
    """BASELINE configs[0] on the CPU, in full: the reference's default path (no -X: one thread) for one O320 field
    -> EFAS 5 km invdist: ScipyInterpolation.__init__ + interpolate (table build), then the apply of one message."""
    import numpy.ma as ma

    from oracle import interp_ref as R
    lat, lon, det = synth.octahedral_grid(320)
    tl, tlo = synth.efas_target()
    z = synth.field_2t(lat, lon)
    t0 = time.perf_counter()
    o = R.BuildOracle(lon, lat, det, z, K_INVDIST, synth.NC_FILL_F64, synth.GRIB_MV, parallel=False, mode='invdist')
    _, w, idx = o.interpolate(tlo, tl)
    t_build = time.perf_counter() - t0
    rec = R.to_record(idx, w)
    v = ma.masked_array(z, mask=synth.missing_mask(lat))
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        with np.errstate(all='ignore'):
            R.apply_as_reference(v, rec['coeffs'], rec['indexes'], K_INVDIST, synth.NC_FILL_F64)
        ts.append(time.perf_counter() - t0)
    return {'kind': 'port', 'cores': 1, 'build_seconds': t_build, 'build_target_pts_per_s': tl.size / t_build,
            'apply_ms_per_message': float(np.median(ts)) * 1e3,
            'sample': 'the whole case: cKDTree(leafsize=30) + self query + k=4 query (workers=1) + numpy weights, '
                      'then np.append + einsum per message (median of 8)'}


def cpu_configs1(row_fraction=8):
    """BASELINE configs[1] on the CPU: O1280 -> GloFAS 0.1 deg nearest, 1/row_fraction of the target rows (-X path:
    cKDTree workers=-1), fixed costs in full; the apply of one message (`v[indexes]`) on the sampled rows, scaled."""
    import numpy.ma as ma

    from oracle import interp_ref as R
    lat, lon, det = synth.octahedral_grid(1280)
    tl, tlo = synth.latlon_target(0.1)
    rows = tl.shape[0] // row_fraction
    o, t_fixed = _o1280_oracle('nearest', 1)
    t0 = time.perf_counter()
    _, w, idx = o.interpolate(tlo[:rows], tl[:rows])
    t_sample = time.perf_counter() - t0
    m_full = tl.size
    est = t_fixed + t_sample * row_fraction
    v = ma.masked_array(synth.field_2t(lat, lon), mask=synth.missing_mask(lat))
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        R.apply_as_reference(v, w, idx, 1, synth.NC_FILL_F64)
        ts.append(time.perf_counter() - t0)
    return {'kind': 'port', 'cores': os.cpu_count(), 'build_seconds_full_estimate': est,
            'build_target_pts_per_s': m_full / est, 'apply_ms_per_message_full_estimate': float(np.median(ts)) * 1e3 * row_fraction,
            'sample': f'{rows * tl.shape[1]} of {m_full} targets; tree + self query {t_fixed:.2f} s in full, target part '
                      f'{t_sample:.2f} s x{row_fraction}; apply = np.append + v[indexes] on the sampled rows x{row_fraction}'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    per_step = 8
    for _ in range(args.warmup):
        pass                                       # numpy needs no warm-up beyond the first call below
    dt_w, _, _ = cpu_apply_reference(1)
    dt, nbytes, sample = cpu_apply_reference(per_step * args.steps)
    value = nbytes / dt / 1e9
    build = cpu_build_reference()
    line = {
        'impl': 'reference', 'metric': 'intertable apply GB/s vs HBM roofline', 'value': value, 'unit': 'GB/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.messages),
        'sample_messages_per_step': per_step,
        'cpu_baseline': {'value': value, 'unit': 'GB/s', 'cores': 1, 'kind': 'port',
                         'sample': f"{sample['what']}; each step = {per_step} of the {args.messages} messages of the "
                                   f"workload (the per-message cost does not depend on the batch: the reference applies "
                                   f"one message per call)"},
        'e2e': {'value': value, 'unit': 'GB/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'messages_per_s': per_step * args.steps / dt,
        'build': build, 'configs0': {'cpu_baseline': cpu_configs0()}, 'configs1': {'cpu_baseline': cpu_configs1()},
        'host': {'cpu_count': os.cpu_count(), 'numpy': np.__version__, 'scipy': __import__('scipy').__version__},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(messages_per_gpu):
    return {'workload': 'BASELINE configs[3] per-GPU slice: intertable apply batch, 51-member ENS x 60 steps '
                        '(3060 O1280 messages) sharded by member over 8 GPUs -> EFAS 5 km (950x1000), invdist k=4, '
                        'with missing values; `build` = configs[2] O1280 -> GloFAS 0.05 deg (7200x3000) invdist k=4',
            'messages_per_gpu': messages_per_gpu, 'n_src': 6599680, 'm_tgt': 950000, 'k': K_INVDIST,
            'mask_policy': 'propagate',
            'l2': 'the resident batch read by each launch (messages_per_gpu x 52.8 MB) is far larger than the '
                  '126 MB L2, so no flush between timed iterations',
            'parallelism': 'messages sharded by rank, no collective (apply); targets sharded + all-gather (build)'}


# =============================================================================== B200 arm
class _Conf:
    def __init__(self):
        self.vars, self.user_vars = {}, {}

    def check_write(self):
        return True

    def dump(self):
        pass


class _Ctx:
    """the slice of pyg2p's execution context the Interpolator reads"""
    is_with_grib_interpolation = False

    def __init__(self, d):
        self._d = d
        self.configuration = type('C', (), {'intertables': _Conf()})()

    def get(self, key, default=None):
        return self._d.get(key, default)


def _timed(torch, fn, reps=5, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    for _ in range(reps):
        fn()
    b_.record()
    torch.cuda.synchronize()
    return a_.elapsed_time(b_) / reps


def pcie_roofline(torch, dev, mb=256):
    """pinned-memory copy bandwidth of this rank's GPU: H2D alone, D2H alone, both at once (GB/s each way).  Under
    torchrun every rank measures at the same time (the callers sit behind a barrier), so the numbers are what the
    host gives N GPUs together - the roofline of the host-buffer (e2e) paths."""
    n = mb << 20
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(h2d, d2h, reps=4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    run(True, True, 1)
    return {'h2d_alone_GBps': run(True, False), 'd2h_alone_GBps': run(False, True), 'both_GBps_each_way': run(True, True),
            'buffer_MB': mb}


def section_small_configs(torch, L, device, dev, hbm_peak, do_cpu):
    """configs[0] (O320 -> EFAS invdist) and configs[1] (O1280 -> GloFAS 0.1 deg nearest): build on this GPU (device
    timed, and through the drop-in class from numpy to numpy), single-message apply with its own roofline."""
    from pyg2p_b200.scipy_interpolation import ScipyInterpolation
    res = {}
    for tag, n_o, target, mode, k in (('configs0', 320, 'efas', 'invdist', K_INVDIST), ('configs1', 1280, 'glofas0.1', 'nearest', 1)):
        lat, lon, det = synth.octahedral_grid(n_o)
        tl, tlo = synth.efas_target() if target == 'efas' else synth.latlon_target(0.1)
        n, m = lat.size, tl.size
        d_lon, d_lat = torch.from_numpy(lon).to(dev), torch.from_numpy(lat).to(dev)
        d_tlo, d_tla = torch.from_numpy(tlo.ravel()).to(dev), torch.from_numpy(tl.ravel()).to(dev)
        mode_code = L.MODE_INVDIST if mode == 'invdist' else L.MODE_NEAREST
        state = {}

        def build_once():
            ix = device.SourceIndex(d_lon, d_lat, det['radius'])
            mub = ix.min_upper_bound(det['Nj'])
            state['o'] = ix.knn(d_tlo, d_tla, k=k, mode=mode_code, mub=mub, want_flags=True)
            state['ix'] = ix

        for _ in range(2):
            build_once()
        torch.cuda.synchronize()
        walls = []
        for _ in range(5):
            t0 = time.perf_counter()
            build_once()
            torch.cuda.synchronize()
            walls.append((time.perf_counter() - t0) * 1e3)
        build_ms = float(np.median(walls))
        o = state['o']
        table = device.DeviceTable(o['idx'], o['coef'], n, grid_shape=tl.shape)
        n_touched = int(torch.unique(table.idx[table.idx < n]).numel())
        # the drop-in class, numpy in -> (result, weights, indexes) numpy out, as the reference is called
        z = synth.field_2t(lat, lon)
        t0 = time.perf_counter()
        si = ScipyInterpolation(lon, lat, det, z, k, synth.NC_FILL_F64, synth.GRIB_MV, parallel=True, mode=mode)
        r_, w_, i_ = si.interpolate(tlo, tl)
        t_api = time.perf_counter() - t0
        assert i_.shape == ((m, k) if k > 1 else (m,))
        # single-message apply (device-resident message, masks propagated by marker)
        n_pad = -(-n // 16) * 16
        msg = torch.zeros((1, n_pad), dtype=torch.float64, device=dev)
        msg[0, :n] = torch.from_numpy(z).to(dev)
        mk = torch.zeros((1, n_pad), dtype=torch.uint8, device=dev)
        mk[0, :n] = torch.from_numpy(synth.missing_mask(lat).view(np.uint8)).to(dev)
        marked = msg.clone()
        device.mark_masked(marked[:, :n], mk[:, :n])
        out1 = torch.empty((1, m), dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > the 126 MB L2

        def apply_once():
            flush.zero_()
            return table.apply(marked[:, :n], synth.NC_FILL_F64, mask_policy=L.MASK_MARKED, out=out1)

        apply_once()
        ms_with_flush = _timed(torch, apply_once, reps=10)
        ms_flush = _timed(torch, lambda: flush.zero_(), reps=10)
        ms_apply = max(ms_with_flush - ms_flush, 1e-4)
        nbytes = algorithmic_bytes_apply(1, n_touched, m, k, True)
        res[tag] = {
            'workload': ('BASELINE configs[0]: one O320 field (421 120 pts) -> EFAS 5 km (950x1000), scipy invdist k=4'
                         if tag == 'configs0' else
                         'BASELINE configs[1]: O1280 (6 599 680 pts) -> GloFAS 0.1 deg (1500x3600), scipy nearest: table build '
                         '(distances stored for every row, none beyond min_upper_bound on a global source) + single-message apply'),
            'build': {'ms': build_ms, 'target_pts_per_s': m / (build_ms * 1e-3), 'reps_ms': [round(x, 3) for x in walls],
                      'includes': 'lat/lon->xyz, cube-sphere index, self k=2 query + max, k-NN, weights; device-resident lat/lon',
                      'host_api_seconds': t_api,
                      'host_api': 'ScipyInterpolation(...).interpolate(lon, lat): numpy in, (result, weights, indexes) numpy out'},
            'apply_single_message': {'ms': ms_apply, 'kernel_path': table.last_path, 'n_touched': n_touched,
                                     'algorithmic_bytes': nbytes,
                                     'roofline': {'bound': 'hbm', 'achieved': nbytes / (ms_apply * 1e-3) / 1e9, 'peak': hbm_peak,
                                                  'unit': 'GB/s', 'frac': nbytes / (ms_apply * 1e-3) / 1e9 / hbm_peak},
                                     'l2': 'a 256 MB buffer is rewritten before every timed apply (its time, measured '
                                           'separately, is subtracted)', 'mask_policy': 'marked (masked cells = mv_out)'},
        }
        if do_cpu:
            res[tag]['cpu_baseline'] = cpu_configs0() if tag == 'configs0' else cpu_configs1()
        del table, o, state, si, msg, mk, marked, out1, flush
        torch.cuda.empty_cache()
    return res


def section_adw(torch, L, device, dev):
    """mode adw (Shepard, k = 11), O320 -> EFAS: k-NN + reference radius + angular weights"""
    from pyg2p_b200.scipy_interpolation import ScipyInterpolation
    lat, lon, det = synth.octahedral_grid(320)
    tl, tlo = synth.efas_target()
    z = synth.field_2t(lat, lon)
    si = ScipyInterpolation(lon, lat, det, z, 11, synth.NC_FILL_F64, synth.GRIB_MV, mode='adw')
    si.build_table(tlo, tl)
    torch.cuda.synchronize()
    walls = []
    for _ in range(3):
        t0 = time.perf_counter()
        si.build_table(tlo, tl)
        torch.cuda.synchronize()
        walls.append((time.perf_counter() - t0) * 1e3)
    raw = si.index.knn(tlo.ravel(), tl.ravel(), k=11, mode=L.MODE_RAW)
    t_lat, t_lon = torch.from_numpy(tl.ravel()).to(dev), torch.from_numpy(tlo.ravel()).to(dev)
    s_lat, s_lon = torch.from_numpy(lat).to(dev), torch.from_numpy(lon).to(dev)
    dist = raw['coef'].clone()
    ms_w = _timed(torch, lambda: (dist.copy_(raw['coef']), device.weights_adw(raw['idx'], dist, t_lat, t_lon, s_lat, s_lon,
                                                                            si.ref_radius)), reps=5)
    ms_c = _timed(torch, lambda: dist.copy_(raw['coef']), reps=5)
    return {'workload': 'SURVEY 8(f)-4: O320 -> EFAS 5 km, mode adw (Shepard angular-distance weighting, k = 11)',
            'build_ms_from_numpy_latlon': float(np.median(walls)), 'targets': int(tl.size),
            'weights_kernel_ms': max(ms_w - ms_c, 1e-4), 'reference_radius_m': si.ref_radius}


def run_b200(args):
    import torch
    import torch.distributed as dist

    from pyg2p_b200 import _lib as L
    from pyg2p_b200 import device, parallel

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: pyg2p_b200 has no CPU fallback (use --impl reference '
                         'for the CPU baseline)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if ws > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL prints its version banner on stdout when the communicator comes up: keep stdout for the one
        # JSON line by pointing fd 1 at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=5):
        return _timed(torch, fn, reps)

    hbm_peak, peak_src = peaks()
    lat, lon, det = synth.octahedral_grid(1280)
    n = lat.size
    tl, tlo = synth.efas_target()
    m = tl.size
    B = args.messages
    do_cpu = rank == 0 and ws == 1 and not args.skip_cpu

    # ---- table for the apply phase, built on this GPU (O1280 -> EFAS invdist)
    ix = device.SourceIndex(lon, lat, det['radius'])
    mub = ix.min_upper_bound(det['Nj'])
    out = ix.knn(tlo, tl, k=K_INVDIST, mode=L.MODE_INVDIST, mub=mub, want_flags=False)
    table = device.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
    touched = torch.unique(table.idx[table.idx < n])
    n_touched = int(touched.numel())
    del touched, out, ix

    # ---- resident batch: this rank's members' messages (synthetic 2t-like fields + GRIB bitmap masks).  `src` /
    # `src_mask` is the form with separate byte masks, `marked` the form the producers of resident fields write (the
    # marker NaN folded in where the bitmap says missing) - the input of the timed step.
    g = torch.Generator(device=dev).manual_seed(20260101 + rank)
    src = torch.empty((B, n), dtype=torch.float64, device=dev)
    lat_d = torch.from_numpy(lat).to(dev)
    base = 288.0 - 45.0 * torch.sin(torch.deg2rad(lat_d)) ** 2
    src.normal_(0.0, 3.0, generator=g)
    src += base
    mask1 = torch.from_numpy(synth.missing_mask(lat).view(np.uint8)).to(dev)
    src_mask = mask1.unsqueeze(0).repeat(B, 1).contiguous()
    src_mask[1::2] |= (torch.rand((B // 2, n), device=dev, generator=g) < 0.003).to(torch.uint8)   # masks differ by message
    src.masked_fill_(src_mask.bool(), synth.GRIB_MV)
    marked = src.clone()
    device.mark_masked(marked, src_mask)
    out_t = torch.empty((B, m), dtype=torch.float64, device=dev)
    out_m = torch.empty((B, m), dtype=torch.uint8, device=dev)
    out_b = torch.empty((B, device.bits_row_bytes(m)), dtype=torch.uint8, device=dev)

    def step():
        table.apply(marked, synth.NC_FILL_F64, mask_policy=L.MASK_MARKED, out=out_t, out_mask=out_b)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.prepare()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = device.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    torch.cuda.nvtx.range_push('timed_apply')       # ncu --nvtx --nvtx-include "timed_apply/" lists exactly these launches
    for a, b_ in ev:
        a.record()
        step()
        b_.record()
    torch.cuda.nvtx.range_pop()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = device.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0][0].elapsed_time(ev[-1][1])
    kern_ms = [a.elapsed_time(b_) for a, b_ in ev]
    total_ms = parallel.max_over_ranks(total_ms, dev)
    bytes_launch = algorithmic_bytes_apply(B, n_touched, m, K_INVDIST, True)
    value = bytes_launch * args.steps * ws / (total_ms * 1e-3) / 1e9
    avg_kernel_ms = float(np.mean(kern_ms))
    achieved = bytes_launch / (avg_kernel_ms * 1e-3) / 1e9
    assert table.last_path == 'window', table.plan_error

    # ---- parity of the timed output: the oracle on whole messages (rank 0, when the CPU legs run), and the other mask
    # forms of the same batch bit for bit on the device (every rank)
    step()
    ref_o, ref_m = torch.empty_like(out_t), torch.empty_like(out_m)
    table.apply(src, synth.NC_FILL_F64, src_mask=src_mask, mask_policy=L.MASK_PROPAGATE, out=ref_o, out_mask=ref_m)
    assert torch.equal(out_t.view(torch.int64), ref_o.view(torch.int64)), 'marked and byte-mask forms differ'
    probe = [0, B // 2, B - 1]
    got_bits = device.unpack_mask_bits(out_b[probe], m)
    assert np.array_equal(got_bits, ref_m[probe].cpu().numpy().astype(bool)), 'bit-packed and byte masks differ'
    parity = {'device_forms_bit_equal': True}
    if do_cpu:
        from oracle import interp_ref as R
        idx_h, w_h = table.indexes_coeffs()
        rec = R.to_record(idx_h, w_h)
        for i in probe[:2]:
            with np.errstate(all='ignore'):
                want, want_mask = R.apply_propagate(src[i].cpu().numpy(), rec['coeffs'], rec['indexes'], K_INVDIST,
                                                    synth.NC_FILL_F64, src_mask[i].cpu().numpy().astype(bool))
            assert np.array_equal(out_t[i].cpu().numpy(), want, equal_nan=True), 'timed apply output differs from the oracle'
            assert np.array_equal(got_bits[probe.index(i)], want_mask), 'timed apply mask differs from the oracle'
        parity['oracle'] = f'messages {probe[:2]} of the timed batch, all {m} rows: values and masks bit-equal to ' \
                           f'oracle.interp_ref.apply_propagate'
        del idx_h, w_h, rec
    del ref_o, ref_m

    # ---- the other mask forms / the generic kernel on the same batch (informational, 5 launches each)
    variants = {}

    def variant(name, fn, masks=True, reps=5):
        ms = timed(fn, reps)
        nb = algorithmic_bytes_apply(B, n_touched, m, K_INVDIST, masks)
        variants[name] = {'ms': ms, 'frac': nb / (ms * 1e-3) / 1e9 / hbm_peak}

    variant('marked fields, no mask array out (masked cells = mv_out; the writer_post path)',
            lambda: table.apply(marked, synth.NC_FILL_F64, mask_policy=L.MASK_MARKED, out=out_t, out_mask=False))
    variant('byte masks in and out (round-1 headline form)',
            lambda: table.apply(src, synth.NC_FILL_F64, src_mask=src_mask, mask_policy=L.MASK_PROPAGATE, out=out_t, out_mask=out_m))
    variant('byte masks in, no mask array out (fill)',
            lambda: table.apply(src, synth.NC_FILL_F64, src_mask=src_mask, mask_policy=L.MASK_FILL, out=out_t))
    variant('as_reference (masks ignored)', lambda: table.apply(src, synth.NC_FILL_F64, out=out_t), masks=False)
    os.environ['G2P_APPLY_PATH'] = 'generic'
    variant('generic gather kernel, byte masks in and out',
            lambda: table.apply(src, synth.NC_FILL_F64, src_mask=src_mask, mask_policy=L.MASK_PROPAGATE, out=out_t, out_mask=out_m),
            reps=2)
    del os.environ['G2P_APPLY_PATH']
    table.apply(src, synth.NC_FILL_F64, src_mask=src_mask, mask_policy=L.MASK_PROPAGATE, out=out_t, out_mask=out_m)
    torch.cuda.synchronize()

    # ---- e2e: through the reference-facing class.  A pyg2p_b200 Interpolator is set up as pyg2p sets one up
    # (execution context, target maps, intertable directory); its first call builds and saves the intertable
    # (untimed).  Each timed step is one `Interpolator.interpolate_batch` call on messages lying in pinned host
    # memory - where the GRIB decoder writes them - and returns host arrays: H2D (the gather kernel reads the
    # touched values over PCIe), apply, D2H are all inside the timed region.
    import tempfile

    from pyg2p_b200.interpolator import Interpolator
    from pyg2p_b200.maps import LatLong

    # one intertable directory for all ranks, as in a real run (rank 0 writes the table, the others wait and load it)
    box = [tempfile.mkdtemp(prefix='g2p_bench_') if rank == 0 else None]
    if ws > 1:
        dist.broadcast_object_list(box, src=0)
    tmpdir = box[0]
    ctx = _Ctx({'interpolation.mode': 'invdist', 'interpolation.create': True, 'interpolation.parallel': True,
                'interpolation.dirs': {'user': tmpdir, 'global': tmpdir}, 'input.file': 'ens_o1280.grib',
                'interpolation.latMap': LatLong(tl, tlo, mv=synth.NC_FILL_F64, identifier='efas5km_laea_1000x950'),
                'interpolation.lonMap': None})
    interp = Interpolator(ctx, synth.GRIB_MV, mask_policy='propagate')
    Be = min(args.e2e_messages, B)
    h_src = interp.pinned_source_buffer(Be, n)
    h_msk = device.pinned_empty((Be, n), torch.uint8)
    h_src.copy_(src[:Be])
    h_msk.copy_(src_mask[:Be])
    h_out = interp.pinned_output_buffer(Be)
    h_om = interp.pinned_output_buffer(Be, torch.uint8)
    grid_id = '0.0$359.929$0$2560$6599680$reduced_gg'
    torch.cuda.synchronize()
    barrier()
    pcie = pcie_roofline(torch, dev)
    barrier()

    def e2e_step():
        return interp.interpolate_batch(lat, lon, h_src, grid_id, det, masks=h_msk, out=h_out, out_mask=h_om)

    first = e2e_step()                     # builds the table, writes tbl_*.npy.gz, creates the host pipeline
    assert isinstance(first, np.ma.MaskedArray) and first.shape == (Be,) + tl.shape
    e2e_step()
    pipe = interp._pipeline(interp._table_for(lat, lon, h_src[0].numpy(), grid_id, det), True)
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.nvtx.range_push('timed_e2e')
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.nvtx.range_pop()
    e1.record()
    barrier()
    e2e_ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev)
    e2e_bytes = algorithmic_bytes_apply(Be, n_touched, m, K_INVDIST, True)
    e2e_value = e2e_bytes * args.steps * ws / (e2e_ms * 1e-3) / 1e9
    assert torch.equal(h_out.to(dev), out_t[:Be]) or torch.allclose(h_out.to(dev), out_t[:Be], equal_nan=True)
    assert torch.equal(h_om.to(dev), out_m[:Be])
    msgs_per_s_rank = Be * args.steps / (e2e_ms * 1e-3)
    d2h_per_msg = pipe.d2h_bytes / args.steps / Be
    e2e = {'value': e2e_value, 'unit': 'GB/s', 'h2d_bytes_per_step': pipe.h2d_bytes // args.steps,
           'd2h_bytes_per_step': pipe.d2h_bytes // args.steps, 'messages_per_step': Be,
           'messages_per_s': Be * args.steps * ws / (e2e_ms * 1e-3),
           'inputs': 'PINNED host buffers, already filled when the clock starts (a decoder writing into '
                     'Interpolator.pinned_source_buffer); `per_message_call` below is the pageable-numpy call pattern of '
                     'the reference',
           'pcie_roofline': dict(pcie, d2h_bytes_per_message=d2h_per_msg,
                                 d2h_bound_messages_per_s_per_gpu=pcie['both_GBps_each_way'] * 1e9 / d2h_per_msg,
                                 frac_of_d2h_bound=msgs_per_s_rank / (pcie['both_GBps_each_way'] * 1e9 / d2h_per_msg),
                                 note='measured on this rank while all ranks copy at once: D2H of the results is the '
                                      'bound of the host-buffer path (8 B per target and message + 1 mask byte)'),
           'api': 'pyg2p_b200.interpolator.Interpolator.interpolate_batch (drop-in for pyg2p Interpolator, mask_policy '
                  'propagate): messages in pinned host buffers -> ' + type(pipe).__name__ + ' (gather kernel reads the '
                  'touched source values over PCIe, windowed apply, D2H) -> numpy MaskedArray views; '
                  'h2d_bytes_per_step counts the touched values pulled over PCIe'}
    # the same call with the writers' per-step post-processing hoisted into the kernel epilogue: fields arrive as
    # the netCDF writer stores them (clone-map mask, value mask, NaN, == mv_output -> fill value) and no mask array
    # comes back (informational; the headline e2e above is the bare interpolation, like the reference arm)
    clone = np.random.default_rng(3).random(tl.shape) < 0.3
    post = interp.writer_post(clone, 9.969209968386869e36, netcdf=True)
    interp.interpolate_batch(lat, lon, h_src, grid_id, det, masks=h_msk, out=h_out, writer_post=post)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    p0.record()
    for _ in range(args.steps):
        ready = interp.interpolate_batch(lat, lon, h_src, grid_id, det, masks=h_msk, out=h_out, writer_post=post)
    p1.record()
    torch.cuda.synchronize()
    post_ms = parallel.max_over_ranks(p0.elapsed_time(p1), dev)
    want = out_t[:Be].clone()
    want[(out_m[:Be] != 0) | torch.isnan(want) | (want == float(interp.mv_output)) |
         torch.from_numpy(clone.ravel()).to(dev)[None, :]] = 9.969209968386869e36
    assert torch.equal(torch.from_numpy(ready.reshape(Be, -1)).to(dev), want)
    e2e['writer_ready_variant'] = {'messages_per_s': Be * args.steps * ws / (post_ms * 1e-3),
                                   'd2h_bytes_per_step': pipe.d2h_bytes // args.steps,
                                   'what': 'interpolate_batch(..., writer_post=...): clone mask / value mask / NaN / '
                                           'mv_output -> fill value in the apply epilogue, no mask array copied back'}
    # ... and handed over as float32, what a PCRaster REAL4 band / netCDF f4 variable stores: half the D2H bytes
    post32 = interp.writer_post(clone, 9.969209968386869e36, netcdf=True, out_dtype=np.float32)
    h_out32 = interp.pinned_output_buffer(Be, torch.float32)
    interp.interpolate_batch(lat, lon, h_src, grid_id, det, masks=h_msk, out=h_out32, writer_post=post32)
    torch.cuda.synchronize()
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    q0.record()
    for _ in range(args.steps):
        ready32 = interp.interpolate_batch(lat, lon, h_src, grid_id, det, masks=h_msk, out=h_out32, writer_post=post32)
    q1.record()
    barrier()
    post32_ms = parallel.max_over_ranks(q0.elapsed_time(q1), dev)
    assert ready32.dtype == np.float32 and torch.equal(torch.from_numpy(ready32.reshape(Be, -1)).to(dev), want.to(torch.float32))
    e2e['writer_ready_float32_variant'] = {
        'messages_per_s': Be * args.steps * ws / (post32_ms * 1e-3), 'd2h_bytes_per_step': pipe.d2h_bytes // args.steps,
        'what': 'writer_post(..., out_dtype=float32): the same fields cast to float32 in the epilogue (round to nearest) - '
                'the precision of the PCRaster REAL4 / netCDF f4 file they are written to - half the D2H bytes'}
    del ready, want, post, ready32, post32, h_out32
    # the reference's own call pattern: one `interpolate_scipy(lat, lon, v, grid_id, details)` per message, v a
    # pageable numpy MaskedArray as ecCodes hands it over
    pm_fields = [np.ma.masked_array(src[i].cpu().numpy(), mask=src_mask[i].cpu().numpy().astype(bool)) for i in range(4)]
    one = interp.interpolate_scipy(lat, lon, pm_fields[0], grid_id, det)
    assert torch.equal(torch.from_numpy(np.ma.getdata(one).reshape(-1)).to(dev), out_t[0])
    barrier()
    pm_t = []
    for i in range(24):
        t0 = time.perf_counter()
        interp.interpolate_scipy(lat, lon, pm_fields[i % 4], grid_id, det)
        pm_t.append((time.perf_counter() - t0) * 1e3)
    pm_t = sorted(pm_t[4:])
    pm_ms = parallel.max_over_ranks(pm_t[len(pm_t) // 2], dev)
    e2e['per_message_call'] = {'ms': pm_ms, 'ms_min_max_rank0': [pm_t[0], pm_t[-1]],
                               'statistic': 'median of 20 calls, max over ranks (all ranks call at the same time)',
                               'messages_per_s': ws * 1e3 / pm_ms,
                               'what': 'Interpolator.interpolate_scipy on one pageable numpy MaskedArray per call '
                                       '(host picks the touched values into a pinned buffer, H2D, apply, D2H, numpy out)'}
    # ... and the batched form of the same pageable inputs: the host picks the touched values of message i+1 while the
    # device works on message i (Interpolator.interpolate_stream)
    if hasattr(interp, 'interpolate_stream'):
        n_stream = 32
        gen_fields = [pm_fields[i % 4] for i in range(n_stream)]
        list(interp.interpolate_stream(lat, lon, iter(gen_fields[:4]), grid_id, det))
        barrier()
        t0 = time.perf_counter()
        n_got = 0
        for res_ in interp.interpolate_stream(lat, lon, iter(gen_fields), grid_id, det):
            n_got += 1
        st_ms = (time.perf_counter() - t0) * 1e3 / n_stream
        assert n_got == n_stream and np.array_equal(np.ma.getdata(res_).reshape(-1), out_t[(n_stream - 1) % 4].cpu().numpy(),
                                                    equal_nan=True)
        st_ms = parallel.max_over_ranks(st_ms, dev)
        e2e['pageable_stream'] = {'ms_per_message': st_ms, 'messages_per_s': ws * 1e3 / st_ms, 'messages': n_stream,
                                  'what': 'Interpolator.interpolate_stream(iterator of pageable MaskedArrays): a host thread '
                                          'picks the touched values of the next messages into a pinned ring while the device '
                                          'applies and returns the previous ones'}
    del pm_fields, one
    import shutil
    barrier()
    if rank == 0:
        shutil.rmtree(tmpdir, ignore_errors=True)
    del interp, first
    del h_src, h_msk, h_out, h_om, pipe, src, src_mask, marked, out_t, out_m, out_b
    torch.cuda.empty_cache()

    # ---- build phase: O1280 -> GloFAS 0.05 deg invdist k=4, targets sharded by whole rows
    build = None
    if not args.skip_build:
        gl, glo = synth.latlon_target(0.05)
        M = gl.size
        lo_i, hi_i = parallel.shard_bounds(M, rank, ws, align=gl.shape[1])
        d_lon = torch.from_numpy(lon).to(dev)
        d_lat = torch.from_numpy(lat).to(dev)
        d_tlo = torch.from_numpy(glo.ravel()[lo_i:hi_i].copy()).to(dev)
        d_tla = torch.from_numpy(gl.ravel()[lo_i:hi_i].copy()).to(dev)
        assembler = parallel.TableAssembler(M, K_INVDIST, gl.shape[1], dev) if ws > 1 else None

        def build_once():
            ixb = device.SourceIndex(d_lon, d_lat, det['radius'])
            mubb = ixb.min_upper_bound(det['Nj'])
            if assembler is None:
                o = ixb.knn(d_tlo, d_tla, k=K_INVDIST, mode=L.MODE_INVDIST, mub=mubb, want_flags=True)
                return ixb, o['flags'], o['idx'], o['coef']
            flags, idx_f, coef_f = assembler.build(ixb, d_tlo, d_tla, L.MODE_INVDIST, mubb)
            return ixb, flags, idx_f, coef_f

        for _ in range(3):      # warm-up: the stream-ordered memory pool reaches its steady size (two indexes alive)
            warm = build_once()
        del warm
        barrier()
        reps = 5
        l0 = device.launch_count()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        torch.cuda.nvtx.range_push('timed_build')
        rep_wall = []
        for _ in range(reps):
            barrier()
            t_rep = time.perf_counter()
            ixb, flags_b, idx_f, coef_f = build_once()
            torch.cuda.synchronize()
            rep_wall.append((time.perf_counter() - t_rep) * 1e3)
        torch.cuda.nvtx.range_pop()
        b1.record()
        barrier()
        # median repetition (each bracketed by a device synchronisation): the stream-ordered memory pool can make
        # a single repetition several times slower when it regrows after the previous index was returned
        build_ms = parallel.max_over_ranks(float(np.median(rep_wall)), dev)
        build_mean_ms = parallel.max_over_ranks(b0.elapsed_time(b1), dev) / reps
        n_tie = int((flags_b & L.FLAG_TIE).ne(0).sum().item())
        build = {'metric': 'intertable build target-pts/s', 'value': M / (build_ms * 1e-3), 'unit': 'target-pts/s',
                 'ms_per_build': build_ms, 'ms_per_build_mean': build_mean_ms, 'reps_ms': [round(x, 3) for x in rep_wall],
                 'statistic': 'median of reps', 'targets': M, 'targets_per_gpu': hi_i - lo_i, 'k': K_INVDIST,
                 'mode': 'invdist', 'scaling': 'strong', 'reps': reps,
                 'includes': 'lat/lon->xyz, cube-sphere index, self k=2 query + max (min_upper_bound), k-NN, '
                             'invdist weights' + (', table shards assembled on every GPU: ' + assembler.how if ws > 1 else ''),
                 'gpu_launches_per_build': (device.launch_count() - l0) // reps,
                 'exact_tie_rows_rank0': n_tie, 'index': ixb.info(),
                 # M*(16 in + k*(8+8) out) + N*(16 in + 24 xyz); the k-NN is instruction / latency
                 # bound (profiles/r01_build.md), the HBM figure is informational
                 'roofline': {'bound': 'hbm', 'unit': 'GB/s', 'peak': hbm_peak,
                              'algorithmic_bytes_per_build': M * (16 + K_INVDIST * 16) + n * 40,
                              'achieved': (M * (16 + K_INVDIST * 16) + n * 40) / (build_ms * 1e-3) / 1e9,
                              'frac': (M * (16 + K_INVDIST * 16) + n * 40) / (build_ms * 1e-3) / 1e9 / hbm_peak},
                 'workload': 'BASELINE configs[2]: O1280 (6 599 680 pts) -> GloFAS 0.05 deg (7200x3000), invdist k=4'}
        if ws > 1:
            # the same without assembling the table: what a build that only WRITES the table needs (every rank
            # compresses and writes the gzip members of its own rows)
            def shard_only():
                ixs = device.SourceIndex(d_lon, d_lat, det['radius'])
                mubs = ixs.min_upper_bound(det['Nj'])
                return ixs.knn(d_tlo, d_tla, k=K_INVDIST, mode=L.MODE_INVDIST, mub=mubs, want_flags=True)
            shard_only()
            walls = []
            for _ in range(reps):
                barrier()
                t_rep = time.perf_counter()
                keep = shard_only()
                torch.cuda.synchronize()
                walls.append((time.perf_counter() - t_rep) * 1e3)
            del keep
            sh_ms = parallel.max_over_ranks(float(np.median(walls)), dev)
            build['shards_only'] = {'ms_per_build': sh_ms, 'value': M / (sh_ms * 1e-3), 'unit': 'target-pts/s',
                                    'what': 'every rank keeps the table rows of its own targets (no exchange): the form a '
                                            'table-writing run needs'}
            # check the assembled table against this rank's own shard
            assert torch.equal(idx_f[:, lo_i:hi_i], shard_only()['idx'])
        # the apply kernel on THIS table (global lat/lon target: rows run along the source rings), 16 messages
        gt = device.DeviceTable(idx_f, coef_f, n, grid_shape=gl.shape)
        gb = 16
        gsrc = torch.randn((gb, n), generator=g, device=dev, dtype=torch.float64)
        gmsk = (torch.rand((gb, n), generator=g, device=dev) < 0.02).to(torch.uint8)
        gmarked = gsrc.clone()
        device.mark_masked(gmarked, gmsk)
        gout = torch.empty((gb, M), dtype=torch.float64, device=dev)
        gom = torch.empty((gb, device.bits_row_bytes(M)), dtype=torch.uint8, device=dev)
        g_touched = int(torch.unique(idx_f[idx_f < n]).numel())
        ms_m = timed(lambda: gt.apply(gmarked, synth.NC_FILL_F64, mask_policy=L.MASK_MARKED, out=gout, out_mask=gom), reps=5)
        ms_p = timed(lambda: gt.apply(gsrc, synth.NC_FILL_F64, out=gout), reps=5)
        build['apply_on_this_table'] = {
            'messages': gb, 'kernel_path': gt.last_path, 'plan': gt.plan_info(),
            'marked_bits': {'ms': ms_m, 'frac_of_hbm_peak': algorithmic_bytes_apply(gb, g_touched, M, K_INVDIST, True) / (ms_m * 1e-3) / 1e9 / hbm_peak},
            'masks_ignored': {'ms': ms_p, 'frac_of_hbm_peak': algorithmic_bytes_apply(gb, g_touched, M, K_INVDIST, False) / (ms_p * 1e-3) / 1e9 / hbm_peak}}
        del gt, gsrc, gmsk, gmarked, gout, gom
        del ixb, flags_b, idx_f, coef_f, assembler
        if ws == 1:
            # the same through the drop-in class, host arrays in -> host record (reference layout) out
            from pyg2p_b200.scipy_interpolation import ScipyInterpolation
            zeros = np.zeros(n)
            t0 = time.perf_counter()
            si = ScipyInterpolation(lon, lat, det, zeros, K_INVDIST, synth.NC_FILL_F64, synth.GRIB_MV, parallel=True,
                                    mode='invdist')
            rec = si.build_table(glo, gl).to_record()
            t_e2e = time.perf_counter() - t0
            build['e2e'] = {'value': M / t_e2e, 'unit': 'target-pts/s', 'seconds': t_e2e,
                            'h2d_bytes': int(lon.nbytes + lat.nbytes + gl.nbytes + glo.nbytes), 'd2h_bytes': int(rec.nbytes),
                            'api': 'ScipyInterpolation(...).build_table(lon, lat).to_record(): pageable numpy lat/lon in, '
                                   'numpy record array {int64 indexes, f8 coeffs} (m, 4) out'}
            # writing the table as the Interpolator does (`.npy.gz`, readable by the reference's loader): every host
            # core compresses a slice (the reference's single-threaded level-9 write takes ~41 s per million k=4 rows)
            import tempfile as _tf
            from pyg2p_b200.interpolator import save_npy_gz_parallel
            with _tf.TemporaryDirectory(prefix='g2p_tbl_') as td:
                tpath = os.path.join(td, 'tbl.npy.gz')
                t0 = time.perf_counter()
                save_npy_gz_parallel(tpath, rec)
                build['e2e']['table_write_seconds'] = time.perf_counter() - t0
                build['e2e']['table_file_bytes'] = os.path.getsize(tpath)
                from pyg2p_b200.interpolator import load_npy_gz_parallel
                t0 = time.perf_counter()
                back = load_npy_gz_parallel(tpath)
                build['e2e']['table_read_seconds'] = time.perf_counter() - t0
                assert back.shape == rec.shape and back[-1].tobytes() == rec[-1].tobytes()
                del back
            del si, rec
        torch.cuda.empty_cache()

    # ---- fused pipeline (BASELINE configs[4], eue_r24-style): members x (steps 0..360 by 6 h) of one GPU's share
    # of the 51-member ensemble -> x*1000 -> 24 h accumulation -> cut-off -> invdist apply -> lapse-rate
    # correction, two launches (g2p_manip on the touched points + windowed apply with the Corrector epilogue)
    pipeline = None
    if not args.skip_pipeline:
        from pyg2p_b200 import manipulation as MP
        members = max(1, 51 // 8)
        steps_h = list(range(0, 361, 6))
        conv = MP.Converter(func='x=x*1000', cut_off=True)
        conv.set_missing_value(synth.GRIB_MV)
        agg = MP.Aggregator(aggr_step=24, aggr_type='accumulation', aggr_halfweights=False, input_step=6,
                            step_type='accum', start_step=0, end_step=360, unit_time=24, mv_grib=synth.GRIB_MV,
                            force_zero_array=False)
        plan = agg.plan(steps_h)
        psrc = torch.empty((members * len(steps_h), n), dtype=torch.float64, device=dev)
        psrc.normal_(0.0, 1.0, generator=g).clamp_(min=0.0).mul_(2e-3)
        psrc = psrc.view(members, len(steps_h), n).cumsum(1).view(members * len(steps_h), n).contiguous()
        pmsk = mask1.unsqueeze(0).repeat(psrc.shape[0], 1).contiguous()
        dem = synth.field_dem(tl, tlo, synth.PCR_MV_F32).astype(np.float32)
        corr = MP.Corrector.from_geopotential(table, synth.field_geopotential(lat, lon), synth.GRIB_MV, synth.NC_FILL_F64,
                                              dem, float(synth.PCR_MV_F32))
        fused = MP.FusedPipeline(table)

        def pipe_step():
            return fused.run(psrc, None, plan, converter=conv, cut_off=True, mv_out=synth.NC_FILL_F64, corrector=corr,
                             members=members)

        def pipe_step_masked():
            return fused.run(psrc, pmsk, plan, converter=conv, cut_off=True, mv_out=synth.NC_FILL_F64, corrector=corr,
                             members=members, mask_policy=L.MASK_MARKED)

        pout = pipe_step()
        n_out = pout.shape[0]
        pipe_step_masked()
        l0 = device.launch_count()
        torch.cuda.nvtx.range_push('timed_pipeline')
        pms = timed(pipe_step, reps=5)
        torch.cuda.nvtx.range_pop()
        pl = (device.launch_count() - l0) // 6
        pms = parallel.max_over_ranks(pms, dev)
        pms_masked = parallel.max_over_ranks(timed(pipe_step_masked, reps=5), dev)
        # algorithmic bytes: 2 input steps x 8 B per touched point + 8 B per target and output, + gem/dem/table once
        pbytes = n_out * (2 * 8 * n_touched + 8 * m) + m * (K_INVDIST * 12 + 17)
        pbytes_masked = pbytes + n_out * 2 * n_touched + n_out * (m // 8)
        ms_single = timed(lambda: fused.run(psrc, None, plan, converter=conv, cut_off=True, mv_out=synth.NC_FILL_F64,
                                            corrector=corr, members=members, fused=True), reps=3)
        pipeline = {'metric': 'fused pipeline output messages/s', 'value': n_out * ws / (pms * 1e-3), 'unit': 'messages/s',
                    'ms_per_pass': pms, 'outputs_per_gpu': n_out, 'inputs_per_gpu': int(psrc.shape[0]),
                    'members_per_gpu': members, 'gpu_launches_per_pass': pl,
                    'ms_per_pass_single_launch_variant': ms_single,
                    'algorithmic_GBps_per_gpu': pbytes / (pms * 1e-3) / 1e9,
                    'frac_of_hbm_peak': pbytes / (pms * 1e-3) / 1e9 / hbm_peak,
                    'masked': {'ms_per_pass': pms_masked, 'frac_of_hbm_peak': pbytes_masked / (pms_masked * 1e-3) / 1e9 / hbm_peak,
                               'what': 'byte masks in, marker NaNs between the two kernels, bit-packed mask out'},
                    'workload': 'BASELINE configs[4]: x*1000 -> 24 h accumulation (steps 0..360 by 6 h) -> cut-off -> '
                                'invdist apply O1280 -> EFAS -> p+gem-dem*0.0065, 6 of 51 members per GPU'}
        # the same chain from HOST buffers with masks (its e2e): the messages of `hm` members lie in pinned host memory, the
        # gather kernel reads the touched values of the steps an output needs over PCIe, results return by D2H
        hm = min(members, args.pipeline_host_members)
        per = len(steps_h)
        hp_src = device.pinned_empty((hm * per, n))
        hp_msk = device.pinned_empty((hm * per, n), torch.uint8)
        hp_src.copy_(psrc[:hm * per])
        hp_msk.copy_(pmsk[:hm * per])
        n_out1 = n_out // members
        hp_out = device.pinned_empty((hm * n_out1, m))
        hp_bits = device.pinned_empty((hm * n_out1, device.bits_row_bytes(m)), torch.uint8)

        def host_pass():
            return fused.run_host(hp_src, hp_msk, plan, converter=conv, cut_off=True, mv_out=synth.NC_FILL_F64,
                                  corrector=corr, members=hm, out=hp_out, out_mask=hp_bits)

        host_pass()
        torch.cuda.synchronize()
        want_o, want_b = fused.run(psrc[:hm * per], pmsk[:hm * per], plan, converter=conv, cut_off=True,
                                   mv_out=synth.NC_FILL_F64, corrector=corr, members=hm, mask_policy=L.MASK_MARKED)
        assert torch.equal(hp_out.to(dev).view(torch.int64), want_o.view(torch.int64)) and torch.equal(hp_bits.to(dev), want_b)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(3):
            host_pass()
        h1.record()
        barrier()
        hms = parallel.max_over_ranks(h0.elapsed_time(h1), dev) / 3
        pipeline['e2e'] = {'value': hm * n_out1 * ws / (hms * 1e-3), 'unit': 'messages/s', 'ms_per_pass': hms,
                           'members_per_pass': hm, 'outputs_per_pass': hm * n_out1, 'inputs_per_pass': hm * per,
                           'h2d_bytes_per_step': int(fused.h2d_bytes), 'd2h_bytes_per_step': int(fused.d2h_bytes),
                           'api': 'manipulation.FusedPipeline.run_host: input messages + byte masks in pinned host memory -> '
                                  'gather/convert/accumulate/cut-off kernel reading the touched values over PCIe -> windowed '
                                  'apply + correction -> D2H of fields and bit-packed masks'}
        del psrc, pmsk, pout, fused, corr, hp_src, hp_msk, hp_out, hp_bits, want_o, want_b
        torch.cuda.empty_cache()

    small = {} if args.skip_small else section_small_configs(torch, L, device, dev, hbm_peak, do_cpu)
    adw = None if args.skip_small else section_adw(torch, L, device, dev)

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only)
    cpu = None
    if do_cpu:
        dt, nbytes, sample = cpu_apply_reference(args.cpu_messages)
        cpu = {'value': nbytes / dt / 1e9, 'unit': 'GB/s', 'cores': 1, 'kind': 'port',
               'sample': sample['what'].replace('of the 383 O1280 messages', 'O1280 messages (4 distinct fields cycled)') +
               f' ({dt:.1f} s); the reference apply is single-threaded numpy',
               'messages_per_s': args.cpu_messages / dt}
        if build is not None:
            build['cpu_baseline'] = cpu_build_reference()

    if rank == 0:
        traffic = None
        tp = os.path.join(REPO, 'profiles', 'apply_traffic.json')
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get('messages') == B and tj.get('mask_form') == 'marked+bits':
                traffic = tj.get('dram_bytes_per_launch')
        line = {
            'metric': 'intertable apply GB/s vs HBM roofline', 'value': value, 'unit': 'GB/s', 'n_gpus': ws,
            'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(B),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                         'frac': achieved / hbm_peak, 'traffic': traffic,
                         'kernel': 'g2p::apply_window_kernel<4, 4, CV, 5, false, false, 1> (marked fields in, bit-packed mask out; five stages handed over through mbarrier rings)',
                         'plan': table.plan_info() or table.plan_error,
                         'algorithmic_bytes_per_launch': bytes_launch, 'avg_launch_ms': avg_kernel_ms,
                         'launch_ms_min_median_max': [float(np.min(kern_ms)), float(np.median(kern_ms)), float(np.max(kern_ms))],
                         'n_touched': n_touched, 'peak_source': peak_src},
            'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
            'messages_per_s': B * args.steps * ws / (total_ms * 1e-3),
            'clocks': clocks, 'wall_s_timed_region': t_wall, 'parity': parity, 'apply_variants': variants, 'build': build,
            'pipeline': pipeline, 'adw': adw,
        }
        line.update(small)
        print(json.dumps(line), flush=True)
    if ws > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--messages', type=int, default=383, help='resident messages per GPU (3060 / 8)')
    ap.add_argument('--e2e-messages', type=int, default=64)
    ap.add_argument('--cpu-messages', type=int, default=400)
    ap.add_argument('--pipeline-host-members', type=int, default=2)
    ap.add_argument('--skip-build', action='store_true')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-pipeline', action='store_true')
    ap.add_argument('--skip-small', action='store_true', help='skip the configs[0] / configs[1] / adw sections')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus != ws and ws == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29531', os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_b200(args)


if __name__ == '__main__':
    sys.exit(main())
