// One membrane position of the ray-tracing model as ONE library call.
//
// Reference: Experiment.computeSampleAndReferenceImages_RT, Experiment.py:407-526 -- the loop over
// the spectrum (:448) with its three or four refraction hops per energy (:463-498), the running
// mean of the reference image (:485-486) and the detector call at every bin close (:501-521).
// The host shim prepares the per-energy scalars in fp64; everything below is kernel launches on
// one stream, so a position costs one Python -> C transition instead of a dozen.
#include "common.cuh"

using namespace paresis;

extern "C" int paresis_rt_run(const paresis_rt_job* job, paresis_stream stream) {
    if (!job || !job->energies_host || job->n_energies < 1 || !job->i_bs || !job->acc_sample || !job->acc_ref ||
        !job->out_sample || !job->out_ref || !job->means ||
        (job->first_point && (!job->acc_propag || !job->acc_white || !job->out_propag || !job->out_white))) {
        set_last_error("paresis_rt_run: incomplete job description");
        return PARESIS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)job->nx * job->ny, nd = (size_t)job->det_x * job->det_y;
    const int margin = 15;   // refractionFileNumba2.py:50
    PARESIS_CUDA(cudaMemsetAsync(job->means, 0, sizeof(double) * job->n_energies, s));
    bool fresh_bin = true;
    int ibin = 0;
    float white = 0.f;
    auto probe = [&](int kind, bool start) {
        if (job->probe == kind && job->probe_start && job->probe_end)
            cudaEventRecord((cudaEvent_t)(start ? job->probe_start : job->probe_end), s);
    };
    for (int e = 0; e < job->n_energies; ++e) {
        const paresis_rt_energy& en = job->energies_host[e];
        if (en.n_hop1 < 1 || en.n_hop2 < 1 || (job->first_point && en.n_propag < 1)) {
            set_last_error("paresis_rt_run: energy %d has an empty hop", e);
            return PARESIS_ERR_ARG;
        }
        if (fresh_bin) {
            PARESIS_CUDA(cudaMemsetAsync(job->acc_sample, 0, sizeof(float) * n, s));
            PARESIS_CUDA(cudaMemsetAsync(job->acc_ref, 0, sizeof(float) * n, s));
            if (job->first_point) PARESIS_CUDA(cudaMemsetAsync(job->acc_propag, 0, sizeof(float) * n, s));
            white = 0.f;
            fresh_bin = false;
        }
        // membrane -> object plane (Experiment.py:463-466)
        PARESIS_CUDA(cudaMemsetAsync(job->i_bs, 0, sizeof(float) * n, s));
        if (e == 0) probe(1, true);
        int rc = paresis_refract_layers(nullptr, en.intensity_membrane, en.hop1, en.n_hop1, job->i_bs, nullptr, nullptr,
                                        nullptr, job->nx, job->ny, margin, job->flag, stream);
        if (e == 0) probe(1, false);
        if (rc) return rc;
        if (e == 0) probe(2, true);
        // object -> detector: sample beam and reference beam in one pass over I_bs (:469-474)
        rc = paresis_refract_layers(job->i_bs, 0.f, en.hop2, en.n_hop2, job->acc_sample, job->acc_ref, nullptr, nullptr,
                                    job->nx, job->ny, margin, job->flag, stream);
        if (e == 0) probe(2, false);
        if (rc) return rc;
        rc = paresis_sum_scaled(job->acc_ref, n, 1.0 / (double)n, job->means + e, stream);   // :485-486 (running)
        if (rc) return rc;
        if (job->first_point) {
            // the sample alone (:490-498); the white field is the incident beam itself
            const bool want_d = job->dx_pad && job->dy_pad && e == job->n_energies - 1;
            if (want_d) {
                const size_t np = (size_t)(job->nx + 2 * margin) * (job->ny + 2 * margin);
                PARESIS_CUDA(cudaMemsetAsync(job->dx_pad, 0, sizeof(float) * np, s));
                PARESIS_CUDA(cudaMemsetAsync(job->dy_pad, 0, sizeof(float) * np, s));
            }
            rc = paresis_refract_layers(nullptr, en.intensity_propag, en.propag, en.n_propag, job->acc_propag, nullptr,
                                        want_d ? job->dx_pad : nullptr, want_d ? job->dy_pad : nullptr, job->nx, job->ny,
                                        margin, job->flag, stream);
            if (rc) return rc;
            white += en.intensity_propag;
        }
        if (en.close_bin) {   // :501-521
            const uint64_t seq = job->sequence + (uint64_t)ibin * 4;
            float* const imgs[4] = {job->acc_sample, job->acc_ref, job->acc_propag, job->acc_white};
            float* const outs[4] = {job->out_sample, job->out_ref, job->out_propag, job->out_white};
            const int count = job->first_point ? 4 : 2;
            if (job->first_point) {
                rc = paresis_fill(job->acc_white, white, n, stream);
                if (rc) return rc;
            }
            for (int k = 0; k < count; ++k) {
                if (ibin == 0 && k == 0) probe(3, true);
                rc = paresis_detect_counts(imgs[k], job->nx, job->ny, job->oversampling, job->det_x, job->det_y,
                                           job->src_kernel, job->src_half, job->psf_kernel, job->psf_half, job->detect_work,
                                           outs[k] + (size_t)ibin * nd, job->noise, job->seed, seq + k, stream);
                if (ibin == 0 && k == 0) probe(3, false);
                if (rc) return rc;
            }
            ++ibin;
            fresh_bin = true;
        }
    }
    return PARESIS_OK;
}
