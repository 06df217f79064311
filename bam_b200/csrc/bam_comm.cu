// bam_comm.cu -- the two exchanges of the path on an NCCL communicator, inside the library:
//   * all-gather of the per-chain moments {count, mean[P], M2[P]} of the sampler, packed into ONE buffer, and the
//     Gelman-Rubin statistic computed redundantly on every rank (chains shard across GPUs);
//   * all-gather of the envelope slices of a prediction whose inputs shard across GPUs;
//   * all-reduce (sum) of device buffers: the per-input histograms of the sample-sharded envelope selection.
// One communicator rank per GPU: either one process per GPU (bamgpu_comm_init_rank: the caller broadcasts the 128-byte
// unique id by whatever means it has -- MPI, a file, torch.distributed in bench.py) or ONE process driving all GPUs of
// the box (bamgpu_comm_init_all = ncclCommInitAll; the bam_gpu driver calls the per-rank functions from one host thread
// per GPU).  NCCL is loaded lazily (dlopen) at the first communicator call, so a single-GPU host never needs it and a
// process that already carries an NCCL (PyTorch's) shares that copy instead of loading a second one.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bam_gpu.h"
#include "bam_handle.h"
#include "bam_host.h"

struct bamgpu_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;
    double* d_send = nullptr;
    double* d_recv = nullptr;
    size_t sendCap = 0, recvCap = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string err;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // a copy that is already in the process (PyTorch's) wins; else the system library
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
            if (!api.lib) api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        for (const char* n : names)
            if (!api.lib) api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) { api.err = "bamgpu_comm: libnccl.so.2 not found (multi-GPU calls need NCCL)"; return; }
#define BAM_SYM(name)                                                                        \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.lib, "nccl" #name));            \
    if (!api.name) api.err = std::string("bamgpu_comm: symbol nccl" #name " missing in libnccl")
        BAM_SYM(GetUniqueId); BAM_SYM(CommInitRank); BAM_SYM(CommInitAll); BAM_SYM(CommDestroy); BAM_SYM(AllGather);
        BAM_SYM(AllReduce); BAM_SYM(GetErrorString); BAM_SYM(GetVersion);
#undef BAM_SYM
    });
    return api;
}

void setmess(char* mess, const std::string& s) {
    if (!mess) return;
    std::strncpy(mess, s.c_str(), BAMGPU_MESS_LEN - 1);
    mess[BAMGPU_MESS_LEN - 1] = 0;
}

#define BAM_NCCL(call)                                                                                       \
    do {                                                                                                     \
        ncclResult_t r_ = (call);                                                                            \
        if (r_ != ncclSuccess) throw bam::CudaError{std::string(#call) + ": " + nccl().GetErrorString(r_)};  \
    } while (0)

template <class F>
int guarded(char* mess, F&& f) {
    if (mess) mess[0] = 0;
    try {
        NcclApi& a = nccl();
        if (!a.err.empty()) throw bam::CudaError{a.err};
        return f();
    } catch (const bam::CudaError& e) {
        setmess(mess, e.what);
        return 1;
    } catch (const std::exception& e) {
        setmess(mess, e.what());
        return 1;
    }
}

void finish_comm(bamgpu_comm* c) {
    BAM_CUDA(cudaSetDevice(c->device));
    BAM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    BAM_CUDA(cudaEventCreate(&c->e0));
    BAM_CUDA(cudaEventCreate(&c->e1));
}

void reserve(bamgpu_comm* c, size_t sendCount, size_t recvCount) {
    if (sendCount > c->sendCap) {
        if (c->d_send) cudaFree(c->d_send);
        BAM_CUDA(cudaMalloc(&c->d_send, sizeof(double) * sendCount));
        c->sendCap = sendCount;
    }
    if (recvCount > c->recvCap) {
        if (c->d_recv) cudaFree(c->d_recv);
        BAM_CUDA(cudaMalloc(&c->d_recv, sizeof(double) * recvCount));
        c->recvCap = recvCount;
    }
}

// [count | mean(P) | M2(P)] per chain, chain-major: one message per rank
__global__ void pack_moments(int K, int P, const double* __restrict__ cnt, const double* __restrict__ mean, const double* __restrict__ m2,
                             double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int W = 1 + 2 * P;
    if (i >= K * W) return;
    const int k = i / W, j = i - k * W;
    out[i] = (j == 0) ? cnt[k] : (j <= P ? mean[(size_t)k * P + (j - 1)] : m2[(size_t)k * P + (j - 1 - P)]);
}

}  // namespace

extern "C" {

int bamgpu_comm_unique_id(char* id128, char* mess) {
    return guarded(mess, [&]() -> int {
        static_assert(sizeof(ncclUniqueId) == BAMGPU_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
        ncclUniqueId id;
        BAM_NCCL(nccl().GetUniqueId(&id));
        std::memcpy(id128, &id, sizeof(id));
        return 0;
    });
}

int bamgpu_comm_init_rank(bamgpu_comm_t** out, int nranks, int rank, const char* id128, int device, char* mess) {
    *out = nullptr;
    return guarded(mess, [&]() -> int {
        if (nranks < 1 || rank < 0 || rank >= nranks) throw bam::CudaError{"bamgpu_comm_init_rank: invalid rank / size"};
        auto* c = new bamgpu_comm();
        c->nranks = nranks; c->rank = rank; c->device = device;
        try {
            BAM_CUDA(cudaSetDevice(device));
            ncclUniqueId id;
            std::memcpy(&id, id128, sizeof(id));
            BAM_NCCL(nccl().CommInitRank(&c->comm, nranks, id, rank));
            finish_comm(c);
        } catch (...) { delete c; throw; }
        *out = c;
        return 0;
    });
}

int bamgpu_comm_init_all(bamgpu_comm_t** out, int ndev, const int* devices, char* mess) {
    for (int i = 0; i < ndev; ++i) out[i] = nullptr;
    return guarded(mess, [&]() -> int {
        if (ndev < 1) throw bam::CudaError{"bamgpu_comm_init_all: no device"};
        std::vector<ncclComm_t> comms(ndev);
        std::vector<int> devs(ndev);
        for (int i = 0; i < ndev; ++i) devs[i] = devices ? devices[i] : i;
        BAM_NCCL(nccl().CommInitAll(comms.data(), ndev, devs.data()));
        for (int i = 0; i < ndev; ++i) {
            auto* c = new bamgpu_comm();
            c->comm = comms[i]; c->nranks = ndev; c->rank = i; c->device = devs[i];
            finish_comm(c);
            out[i] = c;
        }
        return 0;
    });
}

void bamgpu_comm_destroy(bamgpu_comm_t* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->comm && nccl().CommDestroy) nccl().CommDestroy(c->comm);
    if (c->d_send) cudaFree(c->d_send);
    if (c->d_recv) cudaFree(c->d_recv);
    if (c->e0) cudaEventDestroy(c->e0);
    if (c->e1) cudaEventDestroy(c->e1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int bamgpu_comm_info(const bamgpu_comm_t* c, int* nranks, int* rank, int* nccl_version) {
    if (!c) return 1;
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (nccl_version) { *nccl_version = 0; if (nccl().GetVersion) nccl().GetVersion(nccl_version); }
    return 0;
}

int bamgpu_moments_allgather(bamgpu_t* h, bamgpu_comm_t* c, double* rhat, double* count_all, double* mean_all, double* m2_all,
                             double* ms, char* mess) {
    return guarded(mess, [&]() -> int {
        if (!h || !c) throw bam::CudaError{"bamgpu_moments_allgather: null handle"};
        if (!h->haveMoments) throw bam::CudaError{"bamgpu_moments_allgather: run bamgpu_fit first"};
        BAM_CUDA(cudaSetDevice(h->device));
        const int K = h->K, P = h->dp.P, W = 1 + 2 * P;
        const size_t msg = (size_t)K * W;
        reserve(c, msg, msg * c->nranks);
        BAM_CUDA(cudaStreamSynchronize(h->stream));  // the moments are final
        pack_moments<<<(int)((msg + 255) / 256), 256, 0, c->stream>>>(K, P, h->d_mcount, h->d_mmean, h->d_mm2, c->d_send);
        BAM_CUDA(cudaEventRecord(c->e0, c->stream));
        BAM_NCCL(nccl().AllGather(c->d_send, c->d_recv, msg, ncclDouble, c->comm, c->stream));
        BAM_CUDA(cudaEventRecord(c->e1, c->stream));
        std::vector<double> all(msg * c->nranks);
        BAM_CUDA(cudaMemcpyAsync(all.data(), c->d_recv, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, c->stream));
        BAM_CUDA(cudaStreamSynchronize(c->stream));
        if (ms) { float t = 0; cudaEventElapsedTime(&t, c->e0, c->e1); *ms = t; }
        const size_t KT = (size_t)K * c->nranks;  // every rank holds the same number of chains
        std::vector<double> cnt(KT), mean(KT * P), m2(KT * P);
        for (size_t k = 0; k < KT; ++k) {
            const double* r = all.data() + k * W;
            cnt[k] = r[0];
            std::memcpy(&mean[k * P], r + 1, sizeof(double) * P);
            std::memcpy(&m2[k * P], r + 1 + P, sizeof(double) * P);
        }
        if (rhat) bam::rhat_from_moments((int)KT, P, cnt.data(), mean.data(), m2.data(), rhat);
        if (count_all) std::memcpy(count_all, cnt.data(), sizeof(double) * KT);
        if (mean_all) std::memcpy(mean_all, mean.data(), sizeof(double) * KT * P);
        if (m2_all) std::memcpy(m2_all, m2.data(), sizeof(double) * KT * P);
        h->launches += 1;
        return 0;
    });
}

int bamgpu_envelope_allgather(bamgpu_t* h, bamgpu_comm_t* c, int nobs_pred, double* env_all, double* ms, char* mess) {
    return guarded(mess, [&]() -> int {
        if (!h || !c) throw bam::CudaError{"bamgpu_envelope_allgather: null handle"};
        if (!h->d_env) throw bam::CudaError{"bamgpu_envelope_allgather: no resident prediction (bamgpu_predict_resident first)"};
        BAM_CUDA(cudaSetDevice(h->device));
        // rank r holds inputs [r*per, min(nobs, (r+1)*per)) : rows of 7 statistics per output, slice-major on the device
        const int per = (nobs_pred + c->nranks - 1) / c->nranks;
        const int nOut = h->predNout > 0 ? h->predNout : h->dp.nY;
        const int mine = h->predT1 - h->predT0;
        if (h->predT0 != std::min(c->rank * per, nobs_pred) || mine != std::max(0, std::min(nobs_pred, (c->rank + 1) * per) - c->rank * per))
            throw bam::CudaError{"bamgpu_envelope_allgather: this rank's resident prediction is not its slice [rank*per, (rank+1)*per)"};
        const size_t msg = (size_t)per * BAMGPU_NENV * nOut;
        reserve(c, msg, msg * c->nranks);
        BAM_CUDA(cudaStreamSynchronize(h->stream));
        BAM_CUDA(cudaMemsetAsync(c->d_send, 0, sizeof(double) * msg, c->stream));
        // device layout of one output: (t1-t0) x 7 column-major; padded to per x 7 per output in the message
        for (int y = 0; y < nOut; ++y)
            for (int s = 0; s < BAMGPU_NENV; ++s)
                if (mine > 0)
                    BAM_CUDA(cudaMemcpyAsync(c->d_send + ((size_t)y * BAMGPU_NENV + s) * per, h->d_env + ((size_t)y * BAMGPU_NENV + s) * mine,
                                             sizeof(double) * mine, cudaMemcpyDeviceToDevice, c->stream));
        BAM_CUDA(cudaEventRecord(c->e0, c->stream));
        BAM_NCCL(nccl().AllGather(c->d_send, c->d_recv, msg, ncclDouble, c->comm, c->stream));
        BAM_CUDA(cudaEventRecord(c->e1, c->stream));
        std::vector<double> all(msg * c->nranks);
        BAM_CUDA(cudaMemcpyAsync(all.data(), c->d_recv, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, c->stream));
        BAM_CUDA(cudaStreamSynchronize(c->stream));
        if (ms) { float t = 0; cudaEventElapsedTime(&t, c->e0, c->e1); *ms = t; }
        // env_all: nobs_pred x 7 x nOut column-major (the layout of bamgpu_predict's env for t0 = 0, t1 = nobs_pred)
        for (int r = 0; r < c->nranks; ++r) {
            const int t0 = std::min(r * per, nobs_pred), cntR = std::max(0, std::min(nobs_pred, (r + 1) * per) - r * per);
            for (int y = 0; y < nOut; ++y)
                for (int s = 0; s < BAMGPU_NENV; ++s)
                    if (cntR > 0)
                        std::memcpy(env_all + ((size_t)y * BAMGPU_NENV + s) * nobs_pred + t0,
                                    all.data() + (size_t)r * msg + ((size_t)y * BAMGPU_NENV + s) * per, sizeof(double) * cntR);
        }
        return 0;
    });
}

int bamgpu_comm_allreduce_sum(bamgpu_comm_t* c, void* dev_buf, size_t count, int type, void* cuda_stream, char* mess) {
    return guarded(mess, [&]() -> int {
        if (!c) throw bam::CudaError{"bamgpu_comm_allreduce_sum: null communicator"};
        const ncclDataType_t dt = type == 1 ? ncclDouble : (type == 2 ? ncclUint32 : ncclUint64);
        BAM_NCCL(nccl().AllReduce(dev_buf, dev_buf, count, dt, ncclSum, c->comm, cuda_stream ? (cudaStream_t)cuda_stream : c->stream));
        if (!cuda_stream) BAM_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    });
}

int bamgpu_predict_sample_sharded(bamgpu_t* h, bamgpu_comm_t* c, int64_t rep_base, int64_t nrep_total, int doParametric,
                                  const int32_t* doRemnant, uint64_t seed, int t0, int t1, char* mess) {
    return guarded(mess, [&]() -> int {
        if (!h || !c) throw bam::CudaError{"bamgpu_predict_sample_sharded: null handle"};
        if (h->predNrep < 1) throw bam::CudaError{"bamgpu_predict_sample_sharded: no uploaded prediction (bamgpu_predict_upload with this rank's rows of the sample first)"};
        if (t0 < 0 || t1 > h->predNobs || t0 >= t1) throw bam::CudaError{"bamgpu_predict_sample_sharded: bad input range [t0,t1)"};
        if (rep_base < 0 || nrep_total < rep_base + h->predNrep) throw bam::CudaError{"bamgpu_predict_sample_sharded: the shard does not fit the total"};
        BAM_CUDA(cudaSetDevice(h->device));
        h->shardComm = c; h->shardRepBase = rep_base; h->shardNrepTotal = nrep_total;
        try {
            bam::launch_predict(h, doParametric, doRemnant, seed, t0, t1, nullptr);
            BAM_CUDA(cudaStreamSynchronize(h->stream));
        } catch (...) { h->shardComm = nullptr; throw; }
        h->shardComm = nullptr;
        return 0;
    });
}

}  // extern "C"
