"""Device-resident image formation for one membrane position.

This is the B200 execution of ``Experiment.computeSampleAndReferenceImages_RT`` /
``_Fresnel`` (reference Experiment.py:407-526 / :279-405): the per-energy loop runs as a
short sequence of fused CUDA kernels on one stream, the accumulators never leave HBM, and
only the detector images come back to the host.  PyTorch provides device memory and streams;
all arithmetic is in ``libparesis_b200.so`` (C ABI, ``include/paresis_b200.h``).

The host shim (``paresis_b200/CodePython/Experiment.py``) builds a :class:`Scene` from the
XML-derived objects and calls :meth:`ImageFormation.compute_rt` / ``compute_fresnel``.
"""
import time

import numpy as np
import torch

from . import _cabi as abi
from . import hostmath as hm

IMAGES = ("sample", "reference", "propag", "white")


class _PerPosition:
    """Stands for "the membrane map of the position being computed" in a batched scene."""

    def __repr__(self):
        return "PER_POSITION"


PER_POSITION = _PerPosition()


class Layer:
    """One material of an object: a device thickness map (metres) or a uniform thickness."""

    __slots__ = ("map", "uniform", "delta", "beta")

    def __init__(self, thickness, delta, beta):
        # delta / beta: {energy_keV: value}
        if isinstance(thickness, torch.Tensor) or thickness is PER_POSITION:
            self.map, self.uniform = thickness, None
        else:
            self.map, self.uniform = None, float(thickness)
        self.delta, self.beta = delta, beta


class Scene:
    """Everything the per-energy loop reads (products of Experiment.__init__, Experiment.py:81-100).

    ``membrane`` may mix mapped and uniform layers (the support plate is uniform: it only
    attenuates); ``sample`` layers must all be maps."""

    def __init__(self, study_dims, study_pixel_um, oversampling, det_dims, det_pixel_um, psf_sigma,
                 d_source_membrane, d_membrane_object, d_object_detector, mean_shot_count,
                 spectrum, source_size_um, energy_sampling, thresholds, membrane, sample,
                 common_factor=None, plate_factor=None):
        self.study_dims = (int(study_dims[0]), int(study_dims[1]))
        self.study_pixel_um = float(study_pixel_um)
        self.os = int(oversampling)
        self.det_dims = (int(det_dims[0]), int(det_dims[1]))
        self.det_pixel_um = float(det_pixel_um)
        self.psf_sigma = float(psf_sigma)
        self.d1, self.d2, self.d3 = float(d_source_membrane), float(d_membrane_object), float(d_object_detector)
        self.magnification = (self.d1 + self.d3 + self.d2) / (self.d1 + self.d2)
        self.mean_shot_count = float(mean_shot_count)
        self.spectrum = [(float(e), float(w)) for e, w in spectrum]
        self.source_size_um = float(source_size_um)
        self.energy_sampling = float(energy_sampling)
        self.thresholds = list(thresholds)   # already closed by the last energy (Experiment.py:429)
        self.membrane = list(membrane)
        self.sample = list(sample)
        # per-energy scalar factors on the incident intensity: air volume and scintillator
        # efficiency (Experiment.py:452-459); detector protection plate (:478-480, :494-496)
        self.common_factor = common_factor or (lambda e: 1.0)
        self.plate_factor = plate_factor or (lambda e: 1.0)

    def effective_source_fwhm(self):
        """Experiment.py:503 (FWHM in oversampled pixels)."""
        return self.source_size_um * self.d3 / (self.d1 + self.d2) / self.det_pixel_um * self.os


class InsaneValues(Exception):
    pass


def _on_own_device(method):
    """Library calls launch on the CURRENT device: make the engine's device current for the duration of the call."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapper


class ImageFormation:
    """Owns the scratch buffers of one GPU and runs membrane positions one after another."""

    def __init__(self, study_dims, oversampling, det_dims, device=None, seed=None, poisson=True):
        if not torch.cuda.is_available():
            raise RuntimeError("paresis_b200 needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.nx, self.ny = int(study_dims[0]), int(study_dims[1])
        self.os = int(oversampling)
        self.det_x, self.det_y = int(det_dims[0]), int(det_dims[1])
        self.poisson = poisson
        # Detector.py:113 seeds from the wall clock; so do we unless told otherwise
        self.seed = int(np.floor(time.time() * 100 % (2 ** 32 - 1))) if seed is None else int(seed)
        f32 = dict(device=self.device, dtype=torch.float32)
        n = (self.nx, self.ny)
        self.i_bs = torch.zeros(n, **f32)      # invariant of paresis_rt_run: all zero between jobs
        self.i_bs_group = []                   # further such buffers, made when a detector bin holds several energies
        self._i_bs_dirty = False
        self.acc = {k: torch.zeros(n, **f32) for k in IMAGES}
        self.work = torch.empty(abi.detect_work_floats(self.nx, self.ny, self.os, self.det_x, self.det_y), **f32)
        self.expect = torch.empty((self.det_x, self.det_y), **f32)
        self.flag = torch.zeros(1, device=self.device, dtype=torch.int32)
        self.means = torch.zeros(256, device=self.device, dtype=torch.float64)
        self.dx_pad = None
        self.dy_pad = None
        self._kernels = {}
        self._plan = None
        self._vectors = {}
        self._waves = None

    # ------------------------------------------------------------------ helpers
    def _gauss(self, sigma):
        key = round(float(sigma), 12)
        if key not in self._kernels:
            if hm.gaussian_half_width(sigma) == 0:
                self._kernels[key] = None      # 1x1 kernel == identity (round(3*sigma) == 0, Detector.py:212)
            else:
                self._kernels[key] = torch.as_tensor(hm.gaussian_1d(sigma), device=self.device, dtype=torch.float32)
        return self._kernels[key]

    @_on_own_device
    def detect(self, image, fwhm, psf_sigma, sequence, out):
        """Detector.detection (Detector.py:79-119): device image -> ``out`` (device, detector dims)."""
        src = self._gauss(fwhm / 2.355) if fwhm != 0 else None
        psf = self._gauss(psf_sigma) if psf_sigma != 0 else None
        abi.detect_counts(image, self.os, self.det_x, self.det_y, src, psf, self.work, out, self.poisson, self.seed, sequence)

    def check_flag(self):
        """The reference's NaN / 'insane values' guard (refractionFileNumba2.py:81-82), checked lazily."""
        if int(self.flag.item()) & abi.FLAG_NONFINITE:
            self.flag.zero_()
            raise InsaneValues("The calculated intensity refractive includes some nans or insane values")

    @staticmethod
    def _split(layers, energy):
        """-> maps [(tensor, delta, beta)], sum(beta*t) and sum(delta*t) of the uniform layers."""
        maps, ub, ud = [], 0.0, 0.0
        for L in layers:
            d, b = L.delta.get(energy, 0.0), L.beta.get(energy, 0.0)   # a missing energy leaves 0 (Sample.py:303-319)
            if L.map is not None:
                maps.append((L.map, d, b))
            else:
                ub += b * L.uniform
                ud += d * L.uniform
        return maps, ub, ud

    def _new_outputs(self, nbins, first):
        """Detector images of one position, stacked [image][bin][x][y] in ONE device buffer
        (sample, reference; + propag, white at position 0) so they cross PCIe as one copy."""
        count = 4 if first else 2
        stack = torch.empty((count, nbins, self.det_x, self.det_y), device=self.device, dtype=torch.float32)
        out = {k: stack[i] for i, k in enumerate(IMAGES[:count])}
        out["_stack"] = stack
        return out

    def _mean_energy(self, energies, bin_starts=None, sums=None):
        """Experiment.py:485-486, :523.  ``sums`` = per-energy sums of the reference beam (ray tracing:
        formed while it is deposited); otherwise ``self.means`` holds the running means of the
        reference accumulator within each bin (Fresnel)."""
        if sums is not None:
            per = np.asarray(sums, dtype=np.float64) / float(self.nx * self.ny)
        else:
            r = self.means[:len(energies)].cpu().numpy()
            per = r.copy()
            for i in range(1, len(per)):
                if i not in bin_starts:
                    per[i] = r[i] - r[i - 1]
        return float(np.dot(per, energies)), float(per.sum())

    # ------------------------------------------------------------------ ray tracing
    def bins(self, scene):
        """Spectrum indices of each detector bin (closing rule of Experiment.py:501)."""
        out, cur, ibin = [], [], 0
        for ie, (energy, _) in enumerate(scene.spectrum):
            cur.append(ie)
            if energy > scene.thresholds[ibin] - scene.energy_sampling / 2:
                out.append(cur)
                cur, ibin = [], ibin + 1
        if cur:
            out.append(cur)     # energies after the last threshold never reach the detector (as upstream)
        return out

    def _rt_energies(self, scene, indices, closing):
        """ctypes array of per-energy hop scalars for the spectrum entries ``indices``; ``closing`` =
        set of spectrum indices after which the detector runs."""
        s = scene
        g2 = hm.refraction_gradient_scale(s.d2, s.magnification, s.study_pixel_um)
        g3 = hm.refraction_gradient_scale(s.d3, s.magnification, s.study_pixel_um)
        energies = (abi.RtEnergy * len(indices))()
        keep = []
        for slot, ie in enumerate(indices):
            energy, flux = s.spectrum[ie]
            k = hm.wavenumber(energy * 1000)
            i0 = s.mean_shot_count / s.os ** 2 * flux * s.common_factor(energy)      # :438, :451-459
            plate = s.plate_factor(energy)
            mem_maps, mem_ub, _ = self._split(s.membrane, energy)
            smp_maps, smp_ub, _ = self._split(s.sample, energy)
            if smp_ub != 0.0 or not mem_maps or not smp_maps or len(mem_maps) + len(smp_maps) > abi.MAX_LAYERS:
                raise NotImplementedError("scene needs 1+ membrane maps, sample maps only, at most %d maps per hop"
                                          % abi.MAX_LAYERS)
            en = energies[slot]
            # uniform layers and the plate only scale the beam (:463, :478-480)
            en.intensity_membrane = i0 * np.exp(-2 * k * mem_ub) * plate
            en.intensity_propag = i0 * plate
            hop1 = [(t, d * g2, 0.0, 2 * k * b) for t, d, b in mem_maps]
            hop2 = [(t, d * g3, d * g3, 0.0) for t, d, b in mem_maps] + [(t, d * g3, 0.0, 2 * k * b) for t, d, b in smp_maps]
            prop = [(t, d * g3, 0.0, 2 * k * b) for t, d, b in smp_maps]
            for dst, src in ((en.hop1, hop1), (en.hop2, hop2), (en.propag, prop)):
                for m, (t, go, gr, at) in enumerate(src):
                    # a NULL map = the membrane of the position at hand (paresis_rt_run_positions)
                    dst[m].thickness = None if t is PER_POSITION else t.data_ptr()
                    dst[m].grad_obj, dst[m].grad_ref, dst[m].atten = go, gr, at
                    keep.append(t)
            en.n_hop1, en.n_hop2, en.n_propag = len(hop1), len(hop2), len(prop)
            en.close_bin = 1 if ie in closing else 0
        return energies, keep

    def _rt_job(self, scene, energies, first, sequence):
        fwhm = scene.effective_source_fwhm()
        src = self._gauss(fwhm / 2.355) if fwhm != 0 else None
        psf = self._gauss(scene.psf_sigma) if scene.psf_sigma != 0 else None
        if len(energies) > self.means.numel():
            self.means = torch.zeros(len(energies), device=self.device, dtype=torch.float64)
        job = abi.RtJob()
        job.nx, job.ny, job.oversampling, job.det_x, job.det_y = self.nx, self.ny, self.os, self.det_x, self.det_y
        job.first_point, job.n_energies, job.energies_host = int(first), len(energies), energies
        job.i_bs = self.i_bs.data_ptr()
        job.i_bs_dirty = 1 if self._i_bs_dirty else 0
        for k, t in enumerate(self._group_buffers(scene)):
            job.i_bs_group[k] = t.data_ptr()
        job.acc_sample, job.acc_ref = self.acc["sample"].data_ptr(), self.acc["reference"].data_ptr()
        job.acc_propag, job.acc_white = self.acc["propag"].data_ptr(), self.acc["white"].data_ptr()
        job.means, job.detect_work = self.means.data_ptr(), self.work.data_ptr()
        job.src_kernel, job.src_half = (src.data_ptr(), (src.numel() - 1) // 2) if src is not None else (None, 0)
        job.psf_kernel, job.psf_half = (psf.data_ptr(), (psf.numel() - 1) // 2) if psf is not None else (None, 0)
        job.noise, job.seed, job.sequence = int(self.poisson), self.seed, sequence
        job.flag = self.flag.data_ptr()
        return job, (src, psf)

    def _group_buffers(self, scene, owner=None):
        """Extra object-plane buffers so that up to MAX_GROUP energies of a detector bin share one object hop."""
        want = min(abi.MAX_GROUP, max((len(b) for b in self.bins(scene)), default=1)) - 1
        store = self.i_bs_group if owner is None else owner.setdefault("i_bs_group", [])
        while len(store) < want:
            store.append(torch.zeros((self.nx, self.ny), device=self.device, dtype=torch.float32))
        return store[:want]

    @staticmethod
    def sequence(point_num, sequence_base=0):
        """Poisson stream id of a position; image k of bin b draws from sequence + 4b + k."""
        return (sequence_base + point_num) << 16

    @_on_own_device
    def compute_rt(self, scene, point_num, want_displacement=False, sequence_base=0, probe=None, want_mean=True,
                   defer=False):
        """Experiment.computeSampleAndReferenceImages_RT (Experiment.py:407-526) as one library
        call (paresis_rt_run).

        Returns a dict of device tensors [nbins, det_x, det_y]: sample, reference (+ propag, white
        at position 0; absent keys are all-zero images), plus ``mean_energy`` = (sum E*mean,
        sum mean) of Experiment.py:485-486."""
        s = scene
        first = point_num == 0
        bins = self.bins(s)
        nbins, n_e = len(s.thresholds), len(s.spectrum)
        out = self._new_outputs(nbins, first)
        wd = first and want_displacement
        if wd and self.dx_pad is None:
            self.dx_pad = torch.empty((self.nx + 30, self.ny + 30), device=self.device, dtype=torch.float32)
            self.dy_pad = torch.empty_like(self.dx_pad)
        closing = {b[-1] for b in bins[:nbins]}
        energies, keep = self._rt_energies(s, list(range(n_e)), closing)
        job, keep2 = self._rt_job(s, energies, first, self.sequence(point_num, sequence_base))
        job.out_sample, job.out_ref = out["sample"].data_ptr(), out["reference"].data_ptr()
        if first:
            job.out_propag, job.out_white = out["propag"].data_ptr(), out["white"].data_ptr()
        if wd:
            job.dx_pad, job.dy_pad = self.dx_pad.data_ptr(), self.dy_pad.data_ptr()
        launches = n_e * (3 if first else 2) + len(closing) * (2 if first else 1)
        self._run(job, launches, probe)
        if want_mean and defer:
            # no host synchronisation here: the per-energy sums and the status flag travel with the images
            # (one small device buffer); the caller finishes with finish_deferred() once they are on the host
            out["aux"] = torch.cat((self.means[:n_e], self.flag.to(torch.float64)))
            self.flag.zero_()
        elif want_mean:
            out["mean_energy"] = self._mean_energy([e for e, _ in s.spectrum], sums=self.means[:n_e].cpu().numpy())
            self.check_flag()
        return out

    def finish_deferred(self, scene, aux_host):
        """``aux_host`` = host copy of ``out["aux"]`` of a deferred compute_rt: -> mean_energy, raising the
        reference's 'insane values' error if the status flag was set."""
        aux = np.asarray(aux_host, dtype=np.float64)
        if int(aux[-1]) & abi.FLAG_NONFINITE:
            raise InsaneValues("The calculated intensity refractive includes some nans or insane values")
        return self._mean_energy([e for e, _ in scene.spectrum], sums=aux[:-1])

    def _run(self, job, launches, probe=None):
        """paresis_rt_run keeps I_bs all-zero between jobs; a failed call may leave it dirty."""
        self._i_bs_dirty = True
        abi.rt_run(job, launches, probe)
        self._i_bs_dirty = False

    # ------------------------------------------------------------------ many positions, one call
    def _slots(self, n_slots, raster_bytes):
        """Scratch sets of paresis_rt_run_positions: slot 0 is this engine's own buffers."""
        f32 = dict(device=self.device, dtype=torch.float32)
        cache = self.__dict__.setdefault("_slot_cache", [])
        while len(cache) < n_slots:
            k = len(cache)
            cache.append(dict(
                i_bs=self.i_bs if k == 0 else torch.zeros((self.nx, self.ny), **f32),
                acc={name: (self.acc[name] if k == 0 else torch.empty((self.nx, self.ny), **f32)) for name in IMAGES},
                work=None, stream=torch.cuda.Stream(device=self.device), dirty=False))
        for c in cache[:n_slots]:
            if c["work"] is None or c["work"].numel() < raster_bytes:
                c["work"] = torch.empty(raster_bytes, device=self.device, dtype=torch.uint8)
        return cache[:n_slots]

    @_on_own_device
    def compute_rt_positions(self, scene, plan, offsets, points, sequence_base=0, n_slots=2, probe_label=None,
                             probe_events=None, want_means=True, buffers=None, per_launch=0):
        """``len(offsets)`` membrane positions in one library call (paresis_rt_run_positions): per
        position the membrane is rasterised from ``plan`` (geometry.MembranePlan) at ``offsets[p]`` and
        the whole per-energy pipeline runs, positions alternating between ``n_slots`` streams.

        ``per_launch`` > 1: up to that many positions (<= n_slots, <= 8) share ONE launch of the membrane cut, of each
        hop and of the detector (blockIdx.z = position; paresis_rt_job.positions_per_launch) instead of running on
        ``n_slots`` streams -- for grids where one position does not fill the GPU.  Detector bins of one energy only.

        ``scene.membrane`` must hold one ``Layer(PER_POSITION, ...)`` for the rasterised map.
        Returns a dict of device tensors: thickness [P, N, N], sample / reference [P, nbins, dx, dy],
        propag / white [P0, nbins, dx, dy] for the positions with ``points[p] == 0`` (in order),
        sums [P, n_energies] (float64 sums of the reference beam, see paresis_rt_job.means)."""
        s = scene
        n_pos = len(offsets)
        bins = self.bins(s)
        nbins, n_e = len(s.thresholds), len(s.spectrum)
        closing = {b[-1] for b in bins[:nbins]}
        energies, keep = self._rt_energies(s, list(range(n_e)), closing)
        job, keep2 = self._rt_job(s, energies, False, 0)
        job.positions_per_launch = int(per_launch)
        firsts = [p for p in range(n_pos) if points[p] == 0]
        f32 = dict(device=self.device, dtype=torch.float32)
        if buffers is None:
            buffers = dict(
                thickness=torch.empty((n_pos, self.nx, self.ny), **f32),
                images=torch.empty((n_pos, 2, nbins, self.det_x, self.det_y), **f32),
                extra=torch.empty((max(len(firsts), 1), 2, nbins, self.det_x, self.det_y), **f32),
                sums=torch.zeros((n_pos, n_e), device=self.device, dtype=torch.float64))
        field = plan.field()
        membrane = abi.Membrane(plan.table.data_ptr(), plan.table.shape[0], plan.pix, plan.layers, plan.margin,
                                field.data_ptr() if field is not None else None,
                                field.shape[0] if field is not None else 0, field.shape[1] if field is not None else 0)
        raster_bytes = abi.lib.paresis_raster_work_bytes(plan.table.shape[0], plan.layers, self.nx, self.ny)
        slots = self._slots(n_slots, raster_bytes)
        c_slots = (abi.RtSlot * n_slots)()
        for k, c in enumerate(slots):
            sl = c_slots[k]
            sl.i_bs = c["i_bs"].data_ptr()
            sl.acc_sample, sl.acc_ref = c["acc"]["sample"].data_ptr(), c["acc"]["reference"].data_ptr()
            sl.acc_propag, sl.acc_white = c["acc"]["propag"].data_ptr(), c["acc"]["white"].data_ptr()
            sl.raster_work, sl.raster_work_bytes = c["work"].data_ptr(), c["work"].numel()
            sl.stream = c["stream"].cuda_stream
            sl.i_bs_dirty = 1 if (c["dirty"] or (k == 0 and self._i_bs_dirty)) else 0
            for kk, t in enumerate(self._group_buffers(s, owner=None if k == 0 else c)):
                sl.i_bs_group[kk] = t.data_ptr()
            c["dirty"] = True
        self._i_bs_dirty = True
        offs = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64).reshape(n_pos, plan.layers, 2))
        c_pos = (abi.RtPosition * n_pos)()
        launches = 0
        for p in range(n_pos):
            cp = c_pos[p]
            cp.offsets_host = offs[p].ctypes.data_as(abi.ctypes.POINTER(abi.ctypes.c_int64))
            cp.thickness = buffers["thickness"][p].data_ptr()
            cp.out_sample, cp.out_ref = buffers["images"][p, 0].data_ptr(), buffers["images"][p, 1].data_ptr()
            first = points[p] == 0
            if first:
                k = firsts.index(p)
                cp.out_propag, cp.out_white = buffers["extra"][k, 0].data_ptr(), buffers["extra"][k, 1].data_ptr()
            cp.means = buffers["sums"][p].data_ptr() if want_means else None
            cp.sequence = self.sequence(points[p], sequence_base)
            cp.first_point = 1 if first else 0
            if probe_events is not None and probe_events[p] is not None:
                e0, e1 = probe_events[p]
                for e in (e0, e1):
                    if not e.cuda_event:
                        e.record()
                cp.probe_start, cp.probe_end = e0.cuda_event, e1.cuda_event
            launches += (1 if field is not None else 2) + n_e * (3 if first else 2) + len(closing) * (2 if first else 1)
        abi.rt_run_positions(job, membrane, c_pos, c_slots, launches, probe_label if probe_events is not None else None)
        for c in slots:
            c["dirty"] = False
        self._i_bs_dirty = False
        out = dict(thickness=buffers["thickness"], sample=buffers["images"][:, 0], reference=buffers["images"][:, 1],
                   propag=buffers["extra"][:len(firsts), 0], white=buffers["extra"][:len(firsts), 1], sums=buffers["sums"],
                   firsts=firsts, buffers=buffers)
        return out

    @_on_own_device
    def accumulate_rt(self, scene, point_num, indices):
        """The energies ``indices`` (all of ONE detector bin) of a position, without the detector:
        leaves the partial sums in ``self.acc`` and returns the per-energy means of the reference
        beam (device float64 tensor, len(indices)).  Used when a bin's energies are spread over
        several GPUs (shard.py)."""
        first = point_num == 0
        energies, keep = self._rt_energies(scene, list(indices), closing=set())
        job, keep2 = self._rt_job(scene, energies, first, 0)
        dummy = torch.empty(1, device=self.device, dtype=torch.float32)
        job.out_sample = job.out_ref = job.out_propag = job.out_white = dummy.data_ptr()
        self._run(job, len(indices) * (3 if first else 2))
        return self.means[:len(indices)] / float(self.nx * self.ny)

    # ------------------------------------------------------------------ Fresnel
    def _reset(self, n_energies):
        for a in self.acc.values():
            a.zero_()
        if n_energies > self.means.numel():
            self.means = torch.zeros(n_energies, device=self.device, dtype=torch.float64)

    def _close_bin(self, scene, out, ibin, point_num, first, white_sum, sequence_base):
        fwhm = scene.effective_source_fwhm()
        seq = ((sequence_base + point_num) << 16) + ibin * 4
        self.detect(self.acc["sample"], fwhm, scene.psf_sigma, seq + 0, out["sample"][ibin])
        self.detect(self.acc["reference"], fwhm, scene.psf_sigma, seq + 1, out["reference"][ibin])
        if first:
            self.detect(self.acc["propag"], fwhm, scene.psf_sigma, seq + 2, out["propag"][ibin])
            abi.fill(self.acc["white"], white_sum)
            self.detect(self.acc["white"], fwhm, scene.psf_sigma, seq + 3, out["white"][ibin])

    def _fresnel_setup(self):
        if self._plan is None:
            self._plan = abi.FresnelPlan(self.nx, self.ny, 15)
            c64 = dict(device=self.device, dtype=torch.complex64)
            self._waves = [torch.empty((self.nx, self.ny), **c64) for _ in range(2)]

    def _transfer(self, scene, distance, energy, magnification):
        # everything the transfer function depends on (Experiment.py:243-250): the cache outlives a scene
        key = (distance, energy, magnification, tuple(int(v) for v in scene.study_dims), float(scene.study_pixel_um))
        if key not in self._vectors:
            hx, hy, phase = hm.fresnel_vectors(self.nx, self.ny, 15, scene.study_dims, scene.study_pixel_um,
                                               distance, energy, magnification)
            hx, hy = torch.as_tensor(hx, device=self.device), torch.as_tensor(hy, device=self.device)
            # the convolution kernels of this transfer function, prepared once for every position of the scan
            self._vectors[key] = (self._plan.kernel(hx, hy), phase)
        return self._vectors[key]

    @_on_own_device
    def propagate(self, scene, wave_in, distance, energy, magnification, wave_out=None, intensity_acc=None):
        """Experiment.wavePropagation (Experiment.py:219-252) on device fields."""
        self._fresnel_setup()
        kern, phase = self._transfer(scene, distance, energy, magnification)
        # |.|^2 ignores the global phase; a returned field carries it (formed in fp64 on the host)
        self._plan.convolve(wave_in, kern, phase if wave_out is not None else 1.0, wave_out, intensity_acc)

    @_on_own_device
    def compute_fresnel(self, scene, point_num, sequence_base=0):
        """Experiment.computeSampleAndReferenceImages_Fresnel (Experiment.py:279-405)."""
        s = scene
        self._fresnel_setup()
        first = point_num == 0
        out = self._new_outputs(len(s.thresholds), first)
        self._reset(len(s.spectrum))
        w_a, w_b = self._waves
        mag_mem_obj = (s.d1 + s.d2) / s.d1                                            # :340
        ibin, white_sum, bin_starts = 0, 0.0, {0}
        for ie, (energy, flux) in enumerate(s.spectrum):
            k = hm.wavenumber(energy * 1000)
            i0 = s.mean_shot_count / s.os ** 2 * flux * s.common_factor(energy)       # :308, :320-333
            plate = s.plate_factor(energy)
            amp = float(np.sqrt(i0 * plate))   # the plate scales every intensity linearly (:352-356)
            mem_maps, mem_ub, _ = self._split(s.membrane, energy)
            smp_maps, smp_ub, _ = self._split(s.sample, energy)
            if smp_ub != 0.0 or not mem_maps or not smp_maps:
                raise NotImplementedError("scene needs 1+ membrane maps and sample maps only")
            # after the membrane (:338); uniform layers: attenuation folded in, constant phase dropped
            abi.transmit_wave(None, amp * np.exp(-k * mem_ub), [t for t, _, _ in mem_maps],
                              [k * b for _, _, b in mem_maps], [k * d for _, d, _ in mem_maps], w_a)
            # reference beam: membrane -> detector in one hop (:349, :354)
            self.propagate(s, w_a, s.d3 + s.d2, energy, s.magnification, intensity_acc=self.acc["reference"])
            # sample beam: membrane -> object (:340-341), through the sample (:344), -> detector (:348)
            self.propagate(s, w_a, s.d2, energy, mag_mem_obj, wave_out=w_b)
            abi.transmit_wave(w_b, 0.0, [t for t, _, _ in smp_maps], [k * b for _, _, b in smp_maps],
                              [k * d for _, d, _ in smp_maps], w_b)
            self.propagate(s, w_b, s.d3, energy, s.magnification, intensity_acc=self.acc["sample"])
            abi.mean(self.acc["reference"], self.means[ie:ie + 1])
            if first:
                abi.transmit_wave(None, amp, [t for t, _, _ in smp_maps], [k * b for _, _, b in smp_maps],
                                  [k * d for _, d, _ in smp_maps], w_b)               # :365
                self.propagate(s, w_b, s.d3, energy, s.magnification, intensity_acc=self.acc["propag"])
                white_sum += amp * amp                                                # :372-375
            if energy > s.thresholds[ibin] - s.energy_sampling / 2:                   # :378
                self._close_bin(s, out, ibin, point_num, first, white_sum, sequence_base)
                ibin += 1
                if ie + 1 < len(s.spectrum):
                    for a in self.acc.values():
                        a.zero_()
                    white_sum = 0.0
                    bin_starts.add(ie + 1)
        out["mean_energy"] = self._mean_energy([e for e, _ in s.spectrum], bin_starts=bin_starts)
        return out
