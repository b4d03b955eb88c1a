// zkm_api.cu -- the C ABI of include/zkm_b200.h: context, error reporting, host<->device staging.
// Every compute call ends in the CUDA kernels of zkm_ntt.cu / zkm_msm.cu; there is no CPU path.
#include <stdarg.h>
#include <string.h>

#include <thread>

#include "zkm_common.cuh"

namespace zkm {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};
static Shared* g_shared = nullptr;
static std::vector<Context*> g_lanes;
static std::mutex g_init_mu;
static std::mutex g_lane_mu;
static std::condition_variable g_lane_cv;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

Context* ctx() {
    if (g_lanes.empty()) ZKM_FAIL(ZKM_ERR_NOT_INIT, "zkm_init() has not been called (or failed)");
    return g_lanes[0];
}

Context* acquire_lane() {
    std::unique_lock<std::mutex> lk(g_lane_mu);
    if (g_lanes.empty()) ZKM_FAIL(ZKM_ERR_NOT_INIT, "zkm_init() has not been called (or failed)");
    for (;;) {
        for (Context* c : g_lanes) {
            if (!c->busy) {
                c->busy = true;
                return c;
            }
        }
        g_lane_cv.wait(lk);
    }
}

int busy_lane_count() {
    std::lock_guard<std::mutex> lk(g_lane_mu);
    int n = 0;
    for (Context* c : g_lanes) n += c->busy ? 1 : 0;
    return n;
}

void release_lane(Context* c) {
    {
        std::lock_guard<std::mutex> lk(g_lane_mu);
        c->busy = false;
    }
    g_lane_cv.notify_one();
}

template <class Fn>
static int32_t guarded(Fn&& fn) {
    try {
        fn();
        return ZKM_OK;
    } catch (const ZkmError& e) {
        cudaGetLastError();  // clear sticky-free errors so the next call starts clean
        return e.code;
    } catch (const std::exception& e) {
        set_error("internal: %s", e.what());
        return ZKM_ERR_CUDA;
    } catch (...) {
        set_error("internal: unknown exception");
        return ZKM_ERR_CUDA;
    }
}

static void check_curve_group(int curve, int group) {
    if (!curve_known(curve)) ZKM_FAIL(ZKM_ERR_ARG, "unknown curve id %d", curve);
    if (group != 1 && group != 2) ZKM_FAIL(ZKM_ERR_ARG, "group must be 1 or 2, got %d", group);
}

// host -> device through the copy engine; pageable memory is fine (driver stages it)
static void h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes) ZKM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
}

static void msm_host(Context* c, int curve, int group, const void* d_bases, const uint8_t* d_inf, const uint64_t* scalars,
                     size_t n, uint64_t* out_xy, uint8_t* out_inf, const BasesReg* pre = nullptr, size_t pre_offset = 0) {
    if (!out_xy || !out_inf) ZKM_FAIL(ZKM_ERR_ARG, "null output pointer");
    if (n && !scalars) ZKM_FAIL(ZKM_ERR_ARG, "null scalars");
    const int W = coord_words(curve, group);
    c->begin(c->stream);
    const size_t sbytes = (size_t)fr_words(curve) * 8;   // BigInteger256 / BigInteger384
    uint64_t* d_scal = (uint64_t*)c->io_scalars.get((n ? n : 1) * sbytes);
    uint64_t* d_out = (uint64_t*)c->io_out.get((2 * W + 1) * 8);
    h2d(d_scal, scalars, n * sbytes, c->stream);
    msm_run(c, curve, group, d_bases, d_inf, d_scal, n, d_out, c->stream, pre, pre_offset);
    uint64_t* h = (uint64_t*)c->pin_in.get((2 * W + 1) * 8);
    ZKM_CUDA(cudaMemcpyAsync(h, d_out, (2 * W + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    ZKM_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(out_xy, h, 2 * W * 8);
    *out_inf = h[2 * W] ? 1 : 0;
}

static int32_t msm_direct(int32_t curve, int group, const uint64_t* bases_xy, const uint8_t* infinity,
                          const uint64_t* scalars, size_t n, uint64_t* out_xy, uint8_t* out_inf) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        check_curve_group(curve, group);
        if (n && !bases_xy) ZKM_FAIL(ZKM_ERR_ARG, "null bases");
        ZKM_CUDA(cudaSetDevice(c->device));
        const int W = coord_words(curve, group);
        c->begin(c->stream);
        void* d_bases = c->io_bases.get((n ? n : 1) * 2 * W * 8);
        uint8_t* d_inf = nullptr;
        h2d(d_bases, bases_xy, n * 2 * W * 8, c->stream);
        if (infinity) {
            d_inf = (uint8_t*)c->io_inf.get(n ? n : 1);
            h2d(d_inf, infinity, n, c->stream);
        }
        msm_host(c, curve, group, d_bases, d_inf, scalars, n, out_xy, out_inf);
    });
}

}  // namespace zkm

using namespace zkm;

extern "C" {

const char* zkm_last_error(void) { return t_err; }
const char* zkm_version(void) { return "zkmember-b200 0.1 (sm_100a)"; }

int32_t zkm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t zkm_init(int32_t device) {
    return guarded([&] {
        std::lock_guard<std::mutex> lk(g_init_mu);
        if (g_shared) {
            if (g_shared->device != device) ZKM_FAIL(ZKM_ERR_ARG, "already bound to device %d (one process per GPU)", g_shared->device);
            return;
        }
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) {
            cudaGetLastError();
            ZKM_FAIL(ZKM_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        }
        if (device < 0 || device >= count) ZKM_FAIL(ZKM_ERR_ARG, "device %d out of range (0..%d)", device, count - 1);
        ZKM_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        ZKM_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) ZKM_FAIL(ZKM_ERR_CUDA, "device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major, prop.minor);
        // the MSM gathers 64..192-byte base records at random: fetch 32-byte sectors, not 128-byte lines
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
        cudaGetLastError();
        Shared* sh = new Shared();
        sh->device = device;
        sh->sm_count = prop.multiProcessorCount;
        std::vector<Context*> lanes;
        for (int i = 0; i < ZKM_NUM_LANES; i++) {
            Context* c = new Context(sh);
            c->lane_id = i;
            ZKM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
            ZKM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            lanes.push_back(c);
        }
        std::lock_guard<std::mutex> ll(g_lane_mu);
        g_shared = sh;
        g_lanes = lanes;
    });
}

void zkm_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    Shared* sh = g_shared;
    if (!sh) return;
    cudaSetDevice(sh->device);
    cudaDeviceSynchronize();
    std::vector<Context*> lanes;
    {
        std::lock_guard<std::mutex> ll(g_lane_mu);
        lanes.swap(g_lanes);
        g_shared = nullptr;
    }
    for (Context* c : lanes) {
        c->ntt_a.release();
        c->ntt_b.release();
        for (auto& b : c->ws) b.release();
        c->io_scalars.release();
        c->io_bases.release();
        c->io_inf.release();
        c->io_out.release();
        c->pin_in.release();
        c->pin_out.release();
        for (auto& e : c->pev)
            if (e) cudaEventDestroy(e);
        if (c->done_ev) cudaEventDestroy(c->done_ev);
        if (c->sync_ev) cudaEventDestroy(c->sync_ev);
        cudaStreamDestroy(c->stream);
        cudaStreamDestroy(c->copy_stream);
        delete c;
    }
    ntt_release_tables(sh);
    for (auto& kv : sh->bases) {
        cudaFree(kv.second.d_xy);
        if (kv.second.d_inf) cudaFree(kv.second.d_inf);
        if (kv.second.d_table) cudaFree(kv.second.d_table);
    }
    sh->bases.clear();
    delete sh;
}

int32_t zkm_msm_g1(int32_t curve, const uint64_t* bases_xy, const uint8_t* infinity, const uint64_t* scalars, size_t n,
                   uint64_t* out_xy, uint8_t* out_inf) {
    return msm_direct(curve, 1, bases_xy, infinity, scalars, n, out_xy, out_inf);
}
int32_t zkm_msm_g2(int32_t curve, const uint64_t* bases_xy, const uint8_t* infinity, const uint64_t* scalars, size_t n,
                   uint64_t* out_xy, uint8_t* out_inf) {
    return msm_direct(curve, 2, bases_xy, infinity, scalars, n, out_xy, out_inf);
}

static int32_t register_impl(int32_t curve, int32_t group, const uint64_t* xy, const uint8_t* inf, size_t n,
                             uint64_t* handle_out, cudaMemcpyKind kind) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        check_curve_group(curve, group);
        if (!handle_out) ZKM_FAIL(ZKM_ERR_ARG, "null handle_out");
        if (n && !xy) ZKM_FAIL(ZKM_ERR_ARG, "null bases");
        ZKM_CUDA(cudaSetDevice(c->device));
        BasesReg r;
        r.curve = curve;
        r.group = group;
        r.n = n;
        const size_t bytes = n * 2 * coord_words(curve, group) * 8;
        ZKM_CUDA(cudaMalloc(&r.d_xy, bytes ? bytes : 16));
        if (bytes) ZKM_CUDA(cudaMemcpyAsync(r.d_xy, xy, bytes, kind, c->stream));
        if (inf) {
            ZKM_CUDA(cudaMalloc((void**)&r.d_inf, n ? n : 16));
            if (n) ZKM_CUDA(cudaMemcpyAsync(r.d_inf, inf, n, kind, c->stream));
        }
        if (c->opt.msm_precompute) msm_precompute(c, &r, c->stream);
        ZKM_CUDA(cudaStreamSynchronize(c->stream));
        std::lock_guard<std::mutex> rl(c->sh->reg_mu);
        uint64_t h = c->next_handle++;
        c->bases[h] = r;
        *handle_out = h;
    });
}

int32_t zkm_bases_register(int32_t curve, int32_t group, const uint64_t* bases_xy, const uint8_t* infinity, size_t n,
                           uint64_t* handle_out) {
    return register_impl(curve, group, bases_xy, infinity, n, handle_out, cudaMemcpyHostToDevice);
}
int32_t zkm_bases_register_device(int32_t curve, int32_t group, const uint64_t* d_bases_xy, const uint8_t* d_infinity,
                                  size_t n, uint64_t* handle_out) {
    return register_impl(curve, group, d_bases_xy, d_infinity, n, handle_out, cudaMemcpyDeviceToDevice);
}

int32_t zkm_bases_release(uint64_t handle) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        ZKM_CUDA(cudaSetDevice(c->device));
        ZKM_CUDA(cudaDeviceSynchronize());   // no lane may still be reading the bases
        std::lock_guard<std::mutex> rl(c->sh->reg_mu);
        auto it = c->bases.find(handle);
        if (it == c->bases.end()) ZKM_FAIL(ZKM_ERR_HANDLE, "unknown bases handle %llu", (unsigned long long)handle);
        cudaFree(it->second.d_xy);
        if (it->second.d_inf) cudaFree(it->second.d_inf);
        if (it->second.d_table) cudaFree(it->second.d_table);
        c->bases.erase(it);
    });
}

static BasesReg lookup(Context* c, uint64_t handle, size_t offset, size_t n) {
    std::lock_guard<std::mutex> rl(c->sh->reg_mu);
    auto it = c->bases.find(handle);
    if (it == c->bases.end()) ZKM_FAIL(ZKM_ERR_HANDLE, "unknown bases handle %llu", (unsigned long long)handle);
    if (offset > it->second.n || n > it->second.n - offset)
        ZKM_FAIL(ZKM_ERR_HANDLE, "range [%zu, %zu) outside the %zu registered bases", offset, offset + n, it->second.n);
    return it->second;
}

int32_t zkm_msm_registered(uint64_t handle, size_t offset, const uint64_t* scalars, size_t n, uint64_t* out_xy,
                           uint8_t* out_inf) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        ZKM_CUDA(cudaSetDevice(c->device));
        const BasesReg r = lookup(c, handle, offset, n);
        const size_t rec = 2 * coord_words(r.curve, r.group) * 8;
        msm_host(c, r.curve, r.group, (const char*)r.d_xy + offset * rec, r.d_inf ? r.d_inf + offset : nullptr, scalars, n,
                 out_xy, out_inf, &r, offset);
    });
}

int32_t zkm_kzg_commit(uint64_t handle, const uint64_t* coeffs, size_t n, uint64_t* out_xy, uint8_t* out_inf) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        if (!out_xy || !out_inf || (n && !coeffs)) ZKM_FAIL(ZKM_ERR_ARG, "null pointer");
        ZKM_CUDA(cudaSetDevice(c->device));
        // skip_leading_zeros_and_convert_to_bigints: drop the zero coefficients at the low end, keep the offset
        const BasesReg r0 = lookup(c, handle, 0, 0);
        const size_t S = (size_t)fr_words(r0.curve);
        size_t z = 0;
        auto coeff_is_zero = [&](size_t i) {
            uint64_t o = 0;
            for (size_t k = 0; k < S; k++) o |= coeffs[S * i + k];
            return o == 0;
        };
        while (z < n && coeff_is_zero(z)) z++;
        const size_t m = n - z;
        const BasesReg r = lookup(c, handle, z, m);
        if (r.group != 1) ZKM_FAIL(ZKM_ERR_ARG, "KZG powers must be G1 bases");
        const int W = coord_words(r.curve, r.group);
        const size_t rec = 2 * (size_t)W * 8;
        c->begin(c->stream);
        uint64_t* d_scal = (uint64_t*)c->io_scalars.get((m ? m : 1) * S * 8);
        uint64_t* d_out = (uint64_t*)c->io_out.get((2 * W + 1) * 8);
        h2d(d_scal, coeffs + S * z, m * S * 8, c->stream);
        fr_into_repr_run(c, r.curve, d_scal, d_scal, (uint64_t)m, c->stream);   // coeffs.into_repr()
        msm_run(c, r.curve, r.group, (const char*)r.d_xy + z * rec, r.d_inf ? r.d_inf + z : nullptr, d_scal, m, d_out,
                c->stream, &r, z);
        uint64_t* h = (uint64_t*)c->pin_in.get((2 * W + 1) * 8);
        ZKM_CUDA(cudaMemcpyAsync(h, d_out, (2 * W + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        ZKM_CUDA(cudaStreamSynchronize(c->stream));
        memcpy(out_xy, h, 2 * W * 8);
        *out_inf = h[2 * W] ? 1 : 0;
    });
}

int32_t zkm_msm_registered_device(uint64_t handle, size_t offset, const uint64_t* d_scalars, size_t n, uint64_t* d_out,
                                  void* stream) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        if (!d_out || (n && !d_scalars)) ZKM_FAIL(ZKM_ERR_ARG, "null device pointer");
        ZKM_CUDA(cudaSetDevice(c->device));
        const BasesReg r = lookup(c, handle, offset, n);
        const size_t rec = 2 * coord_words(r.curve, r.group) * 8;
        cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
        c->begin(s);
        msm_run(c, r.curve, r.group, (const char*)r.d_xy + offset * rec, r.d_inf ? r.d_inf + offset : nullptr, d_scalars, n,
                d_out, s, &r, offset);
        c->end(s);
    });
}

// `count` MSMs over registered bases issued CONCURRENTLY: one host thread + one lane + the lane's own
// stream per item, all ordered after the current point of `stream` and joined back into it.  This is the
// create_proof pattern (four G1 MSMs and the G2 one, ark-groth16 0.3.0 src/prover.rs): upstream runs them
// back to back; here the G2 MSM and the serial tails of the G1 ones overlap.
int32_t zkm_msm_batch_registered_device(int32_t count, const uint64_t* handles, const size_t* offsets,
                                        const uint64_t* const* d_scalars, const size_t* n, uint64_t* const* d_outs,
                                        void* stream) {
    return guarded([&] {
        if (count < 0 || count > ZKM_NUM_LANES - 2) ZKM_FAIL(ZKM_ERR_ARG, "count must be 0..%d", ZKM_NUM_LANES - 2);
        if (count && (!handles || !offsets || !d_scalars || !n || !d_outs)) ZKM_FAIL(ZKM_ERR_ARG, "null argument array");
        Context* c0 = ctx();
        ZKM_CUDA(cudaSetDevice(c0->device));
        cudaStream_t caller = stream ? (cudaStream_t)stream : c0->stream;
        cudaEvent_t ev_in;
        ZKM_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
        ZKM_CUDA(cudaEventRecord(ev_in, caller));
        std::vector<int32_t> rc((size_t)count, ZKM_OK);
        std::vector<std::string> msg((size_t)count);
        std::vector<cudaEvent_t> ev_out((size_t)count, nullptr);
        std::vector<std::thread> th;
        for (int i = 0; i < count; i++) {
            th.emplace_back([&, i] {
                rc[i] = guarded([&] {
                    LaneGuard lane;
                    Context* c = lane.c;
                    ZKM_CUDA(cudaSetDevice(c->device));
                    const BasesReg r = lookup(c, handles[i], offsets[i], n[i]);
                    const size_t rec = 2 * (size_t)coord_words(r.curve, r.group) * 8;
                    cudaStream_t s = c->stream;
                    c->begin(s);
                    ZKM_CUDA(cudaStreamWaitEvent(s, ev_in, 0));
                    msm_run(c, r.curve, r.group, (const char*)r.d_xy + offsets[i] * rec, r.d_inf ? r.d_inf + offsets[i] : nullptr,
                            d_scalars[i], n[i], d_outs[i], s, &r, offsets[i]);
                    c->end(s);
                    ZKM_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
                    ZKM_CUDA(cudaEventRecord(ev_out[i], s));
                });
                if (rc[i] != ZKM_OK) msg[i] = zkm_last_error();
            });
        }
        for (auto& t : th) t.join();
        int32_t first = ZKM_OK;
        for (int i = 0; i < count; i++) {
            if (ev_out[i]) {
                cudaStreamWaitEvent(caller, ev_out[i], 0);
                cudaEventDestroy(ev_out[i]);
            }
            if (rc[i] != ZKM_OK && first == ZKM_OK) {
                first = rc[i];
                set_error("item %d: %s", i, msg[i].c_str());
            }
        }
        cudaEventDestroy(ev_in);
        if (first != ZKM_OK) throw ZkmError{first};
    });
}

int32_t zkm_points_sum_device(int32_t curve, int32_t group, const uint64_t* d_points, size_t m, uint64_t* d_out,
                              void* stream) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        check_curve_group(curve, group);
        if (!d_out || (m && !d_points)) ZKM_FAIL(ZKM_ERR_ARG, "null device pointer");
        ZKM_CUDA(cudaSetDevice(c->device));
        points_sum_run(c, curve, group, d_points, m, d_out, stream ? (cudaStream_t)stream : c->stream);
    });
}

int32_t zkm_ntt(int32_t curve, uint64_t* data, uint32_t log_n, int32_t inverse, int32_t coset) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        if (!curve_known(curve)) ZKM_FAIL(ZKM_ERR_ARG, "unknown curve id %d", curve);
        if (!data) ZKM_FAIL(ZKM_ERR_ARG, "null data");
        const int adicity = fr_two_adicity(curve);
        if ((int)log_n > adicity) ZKM_FAIL(ZKM_ERR_DOMAIN, "log_n %u exceeds the two-adicity %d of Fr", log_n, adicity);
        if (log_n > 30) ZKM_FAIL(ZKM_ERR_ARG, "log_n %u: domains above 2^30 are not supported by this build", log_n);
        ZKM_CUDA(cudaSetDevice(c->device));
        const size_t bytes = ((size_t)fr_words(curve) * 8) << log_n;
        c->begin(c->stream);
        uint64_t* d_a = (uint64_t*)c->io_scalars.get(bytes);
        uint64_t* d_b = (uint64_t*)c->ntt_b.get(bytes);
        h2d(d_a, data, bytes, c->stream);
        ntt_run(c, curve, d_a, d_b, log_n, inverse != 0, coset != 0, c->stream);
        ZKM_CUDA(cudaMemcpyAsync(data, d_b, bytes, cudaMemcpyDeviceToHost, c->stream));
        ZKM_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int32_t zkm_ntt_device(int32_t curve, const uint64_t* d_in, uint64_t* d_out, uint32_t log_n, int32_t inverse,
                       int32_t coset, void* stream) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        if (!d_in || !d_out) ZKM_FAIL(ZKM_ERR_ARG, "null device pointer");
        ZKM_CUDA(cudaSetDevice(c->device));
        cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
        c->begin(s);
        ntt_run(c, curve, d_in, d_out, log_n, inverse != 0, coset != 0, s);
        c->end(s);
    });
}

int32_t zkm_fr_into_repr_device(int32_t curve, const uint64_t* d_in, uint64_t* d_out, size_t n, void* stream) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        if (n && (!d_in || !d_out)) ZKM_FAIL(ZKM_ERR_ARG, "null device pointer");
        ZKM_CUDA(cudaSetDevice(c->device));
        fr_into_repr_run(c, curve, d_in, d_out, (uint64_t)n, stream ? (cudaStream_t)stream : c->stream);
    });
}

int32_t zkm_witness_map_device(int32_t curve, uint64_t* d_a, uint64_t* d_b, uint64_t* d_c, uint32_t log_n, uint64_t* d_h,
                               void* stream) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        if (!d_a || !d_b || !d_c || !d_h) ZKM_FAIL(ZKM_ERR_ARG, "null device pointer");
        ZKM_CUDA(cudaSetDevice(c->device));
        cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
        c->begin(s);
        witness_map_run(c, curve, d_a, d_b, d_c, log_n, d_h, s);
        c->end(s);
    });
}

int32_t zkm_witness_map(int32_t curve, const uint64_t* a, const uint64_t* b, const uint64_t* cc, uint32_t log_n,
                        uint64_t* h_out) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        if (!a || !b || !cc || !h_out) ZKM_FAIL(ZKM_ERR_ARG, "null pointer");
        if (!curve_known(curve)) ZKM_FAIL(ZKM_ERR_ARG, "unknown curve id %d", curve);
        const int adicity = fr_two_adicity(curve);
        if ((int)log_n > adicity) ZKM_FAIL(ZKM_ERR_DOMAIN, "log_n %u exceeds the two-adicity %d of Fr", log_n, adicity);
        if (log_n > 28) ZKM_FAIL(ZKM_ERR_ARG, "log_n %u: witness maps above 2^28 are not supported by this build", log_n);
        ZKM_CUDA(cudaSetDevice(c->device));
        const size_t bytes = ((size_t)fr_words(curve) * 8) << log_n;
        c->begin(c->stream);
        char* d = (char*)c->io_scalars.get(4 * bytes);
        h2d(d, a, bytes, c->stream);
        h2d(d + bytes, b, bytes, c->stream);
        h2d(d + 2 * bytes, cc, bytes, c->stream);
        witness_map_run(c, curve, (uint64_t*)d, (uint64_t*)(d + bytes), (uint64_t*)(d + 2 * bytes), log_n,
                        (uint64_t*)(d + 3 * bytes), c->stream);
        ZKM_CUDA(cudaMemcpyAsync(h_out, d + 3 * bytes, bytes, cudaMemcpyDeviceToHost, c->stream));
        ZKM_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int32_t zkm_domain_constants(int32_t curve, uint32_t log_n, uint64_t* out5x4) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        if (!out5x4) ZKM_FAIL(ZKM_ERR_ARG, "null output pointer");
        ZKM_CUDA(cudaSetDevice(c->device));
        c->begin(c->stream);
        ntt_domain_constants(c, curve, log_n, out5x4);
    });
}

int32_t zkm_set_option(const char* key, int64_t value) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        if (!key) ZKM_FAIL(ZKM_ERR_ARG, "null key");
        if (!strcmp(key, "msm_window_bits")) {
            if (value < 0 || value > 24) ZKM_FAIL(ZKM_ERR_ARG, "msm_window_bits must be 0 (auto) or 2..24");
            c->opt.msm_window_bits = (int)value;
        } else if (!strcmp(key, "msm_chunk")) {
            if (value < 0 || value > 1024) ZKM_FAIL(ZKM_ERR_ARG, "msm_chunk must be 0 (auto) or 1..1024");
            c->opt.msm_chunk = (int)value;
        } else if (!strcmp(key, "ntt_max_radix_log")) {
            if (value < 6 || value > 12) ZKM_FAIL(ZKM_ERR_ARG, "ntt_max_radix_log must be 6..12");
            c->opt.ntt_max_radix_log = (int)value;
        } else if (!strcmp(key, "profile")) {
            c->opt.profile = value ? 1 : 0;
        } else if (!strcmp(key, "msm_affine_levels")) {
            if (value < -1 || value > 24) ZKM_FAIL(ZKM_ERR_ARG, "msm_affine_levels must be -1 (auto) or 0..24");
            c->opt.msm_affine_levels = (int)value;
        } else if (!strcmp(key, "msm_pair_m")) {
            if (value < 2 || value > 4096) ZKM_FAIL(ZKM_ERR_ARG, "msm_pair_m must be 2..4096");
            c->opt.msm_pair_m = (int)value;
        } else if (!strcmp(key, "msm_pair_m2")) {
            if (value < 2 || value > 4096) ZKM_FAIL(ZKM_ERR_ARG, "msm_pair_m2 must be 2..4096");
            c->opt.msm_pair_m2 = (int)value;
        } else if (!strcmp(key, "msm_prefetch_fwd") || !strcmp(key, "msm_prefetch_bwd")) {
            if (value < 0 || value > 64) ZKM_FAIL(ZKM_ERR_ARG, "%s must be 0 (off) .. 64 pairs ahead", key);
            (key[13] == 'f' ? c->opt.msm_prefetch_fwd : c->opt.msm_prefetch_bwd) = (int)value;
        } else if (!strcmp(key, "msm_xarr")) {
            c->opt.msm_xarr = value ? 1 : 0;
        } else if (!strcmp(key, "msm_fold")) {
            if (value != 0 && (value < 2 || value > 1024)) ZKM_FAIL(ZKM_ERR_ARG, "msm_fold must be 0 (auto) or 2..1024");
            c->opt.msm_fold = (int)value;
        } else if (!strcmp(key, "msm_precompute")) {
            c->opt.msm_precompute = value ? 1 : 0;
        } else {
            ZKM_FAIL(ZKM_ERR_ARG, "unknown option '%s'", key);
        }
    });
}

int32_t zkm_profile_last_msm(double* ms_out6) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        double* ms_out5 = ms_out6;
        if (!ms_out6) ZKM_FAIL(ZKM_ERR_ARG, "null output pointer");
        if (!c->pev_valid) ZKM_FAIL(ZKM_ERR_ARG, "no profiled MSM yet (set option \"profile\" to 1 first)");
        ZKM_CUDA(cudaSetDevice(c->device));
        ZKM_CUDA(cudaEventSynchronize(c->pev[6]));
        for (int i = 0; i < 6; i++) {
            float ms = 0.f;
            ZKM_CUDA(cudaEventElapsedTime(&ms, c->pev[i], c->pev[i + 1]));
            ms_out5[i] = ms;
        }
    });
}

uint64_t zkm_launch_count(int32_t reset) {
    uint64_t v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

int32_t zkm_msm_window_bits(int32_t curve, int32_t group, size_t n) {
    if (g_shared && g_shared->opt.msm_window_bits > 0) return g_shared->opt.msm_window_bits;
    return msm_auto_window_bits(curve, group, n);
}

int32_t zkm_testgen_progression_device(int32_t curve, int32_t group, uint64_t a0, uint64_t d, size_t n,
                                       uint64_t* d_bases_xy, void* stream) {
    return guarded([&] {
        LaneGuard lane;
        Context* c = lane.c;
        check_curve_group(curve, group);
        if (n && !d_bases_xy) ZKM_FAIL(ZKM_ERR_ARG, "null device pointer");
        ZKM_CUDA(cudaSetDevice(c->device));
        testgen_progression(c, curve, group, a0, d, n, d_bases_xy, stream ? (cudaStream_t)stream : c->stream);
    });
}

}  // extern "C"
