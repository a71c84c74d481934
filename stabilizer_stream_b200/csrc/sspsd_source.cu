// sspsd_source.cu -- device-resident synthetic sources behind the C ABI (include/sspsd.h).
//
// Mirrors Source::new / Source::get for Data::Noise and Data::Dsm (src/source.rs:66-79, 104-134):
// the generator state persists between calls, so the produced stream does not depend on how it is
// cut into calls (the reference cuts it into 4096-sample blocks).
#include <cmath>
#include <new>

#include "sspsd_cascade.cuh"
#include "sspsd_source_kernel.cuh"

namespace sspsd {

class Source {
public:
    ~Source()
    {
        DeviceGuard g(device_);
        cudaFree(d_state_);
        cudaFree(d_elems_);
        cudaFree(d_init_);
        cudaFree(d_buf_);
        if (own_stream_ && stream_) cudaStreamDestroy(stream_);
    }

    int init(int kind, int64_t param, uint64_t seed, int device, void* stream)
    {
        kind_ = kind;
        device_ = device;
        seed_ = seed;
        if (kind == SSPSD_SOURCE_NOISE) {
            diff_ = param > 0;  // source.rs:70
            const int64_t o = param < 0 ? -param : param;
            if (o > (diff_ ? SRC_MAX_DIFF : SRC_MAX_ORDER)) {
                set_error("noise exponent outside the device range (-4 ..= 8)");
                return SSPSD_EUNIMPLEMENTED;
            }
            order_ = (int)o;
        } else if (kind == SSPSD_SOURCE_DSM) {
            if (param < 0 || param > 0xffffffffll) {
                set_error("dsm tuning word must fit u32");
                return SSPSD_EINVAL;
            }
            ftw_ = (uint32_t)param;
        } else {
            set_error("unknown source kind");
            return SSPSD_EINVAL;
        }
        DeviceGuard g(device_);
        if (!g.ok) return SSPSD_ECUDA;
        if (stream) {
            stream_ = stream == (void*)1 ? cudaStreamLegacy : (cudaStream_t)stream;
        } else {
            SSPSD_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
            own_stream_ = true;
        }
        SSPSD_CUDA(cudaMalloc(&d_state_, 8 * sizeof(double)));
        return reset();
    }

    int reset()
    {
        DeviceGuard g(device_);
        if (!g.ok) return SSPSD_ECUDA;
        if (last_stream_) SSPSD_CUDA(cudaStreamSynchronize(last_stream_));
        SSPSD_CUDA(cudaMemsetAsync(d_state_, 0, 8 * sizeof(double), stream_));
        SSPSD_CUDA(cudaStreamSynchronize(stream_));
        last_stream_ = nullptr;
        pos_ = 0;
        return SSPSD_OK;
    }

    // produce the next n samples of the stream into device memory, in order on stream s
    int generate(float* d_out, size_t n, cudaStream_t s)
    {
        DeviceGuard g(device_);
        if (!g.ok) return SSPSD_ECUDA;
        // the recurrence state lives on the device: order this call after the previous one's stream
        if (last_stream_ && last_stream_ != s) SSPSD_CUDA(cudaStreamSynchronize(last_stream_));
        last_stream_ = s;
        const size_t max_call = (size_t)1 << 28;
        while (n) {
            const size_t c = n < max_call ? n : max_call;
            int rc = generate_call(d_out, c, s);
            if (rc) return rc;
            d_out += c;
            n -= c;
        }
        return SSPSD_OK;
    }

    // scratch buffer for sspsd_cascade_process_source
    int scratch(size_t n, float** out)
    {
        if (n > buf_cap_) {
            DeviceGuard g(device_);
            if (!g.ok) return SSPSD_ECUDA;
            // the old buffer may still be read by the cascade the last call fed (its stream, not ours)
            SSPSD_CUDA(cudaStreamSynchronize(stream_));
            if (last_stream_) SSPSD_CUDA(cudaStreamSynchronize(last_stream_));
            SSPSD_CUDA(cudaFree(d_buf_));
            d_buf_ = nullptr;
            buf_cap_ = 0;
            SSPSD_CUDA(cudaMalloc(&d_buf_, n * sizeof(float)));
            buf_cap_ = n;
        }
        *out = d_buf_;
        return SSPSD_OK;
    }

    uint64_t position() const { return pos_; }
    int seek(uint64_t pos)
    {
        // only the counter-based part of the stream can be entered anywhere: white noise and its differences
        if (kind_ == SSPSD_SOURCE_DSM || (order_ > 0 && !diff_)) {
            set_error("seek: integrated noise and the modulator carry state");
            return SSPSD_EUNIMPLEMENTED;
        }
        pos_ = pos;
        return SSPSD_OK;
    }
    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }

private:
    template <typename Mo>
    int scan_generate(SourceParams& p, cudaStream_t s)
    {
        using E = Aff<typename Mo::T, Mo::K>;
        p.nblocks = (unsigned int)((p.n + SRC_SPB - 1) / SRC_SPB);
        if (p.nblocks > blocks_cap_) {
            SSPSD_CUDA(cudaStreamSynchronize(s));
            SSPSD_CUDA(cudaFree(d_elems_));
            SSPSD_CUDA(cudaFree(d_init_));
            d_elems_ = d_init_ = nullptr;
            blocks_cap_ = 0;
            // sized for the largest map (K = 4 doubles) so one allocation serves every model
            SSPSD_CUDA(cudaMalloc(&d_elems_, (size_t)p.nblocks * 8 * sizeof(double)));
            SSPSD_CUDA(cudaMalloc(&d_init_, (size_t)p.nblocks * 4 * sizeof(double)));
            blocks_cap_ = p.nblocks;
        }
        static_assert(sizeof(E) <= 8 * sizeof(double), "summary size");
        p.block_elems = d_elems_;
        p.block_init = d_init_;
        source_reduce_kernel<Mo><<<p.nblocks, SRC_NT, 0, s>>>(p);
        source_scan_kernel<Mo><<<1, SRC_NT, 0, s>>>(p);
        source_apply_kernel<Mo><<<p.nblocks, SRC_NT, 0, s>>>(p);
        SSPSD_CUDA(cudaGetLastError());
        return SSPSD_OK;
    }

    int generate_call(float* d_out, size_t n, cudaStream_t s)
    {
        SourceParams p{};
        p.pos = pos_;
        p.n = n;
        p.key0 = (uint32_t)seed_;
        p.key1 = (uint32_t)(seed_ >> 32);
        p.ftw = ftw_;
        p.order = order_;
        p.scale = sqrtf(12.0f);
        p.out = d_out;
        p.state = d_state_;
        int rc = SSPSD_OK;
        if (kind_ == SSPSD_SOURCE_DSM) {
            rc = scan_generate<MashModel>(p, s);
        } else if (diff_ || order_ == 0) {
            const unsigned long long nq = ((p.pos + p.n + 3) >> 2) - (p.pos >> 2);
            unsigned long long blocks = (nq + 255) / 256;
            if (blocks > (1ull << 20)) blocks = 1ull << 20;
            source_diff_kernel<<<(unsigned int)blocks, 256, 0, s>>>(p);
            SSPSD_CUDA(cudaGetLastError());
        } else {
            switch (order_) {
            case 1: rc = scan_generate<IntegratorModel<1>>(p, s); break;
            case 2: rc = scan_generate<IntegratorModel<2>>(p, s); break;
            case 3: rc = scan_generate<IntegratorModel<3>>(p, s); break;
            default: rc = scan_generate<IntegratorModel<4>>(p, s); break;
            }
        }
        if (rc) return rc;
        pos_ += n;
        return SSPSD_OK;
    }

    int kind_ = 0, device_ = 0, order_ = 0;
    bool diff_ = false, own_stream_ = false;
    uint64_t seed_ = 0, pos_ = 0;
    uint32_t ftw_ = 0;
    cudaStream_t stream_ = nullptr, last_stream_ = nullptr;
    void* d_state_ = nullptr;
    void* d_elems_ = nullptr;
    void* d_init_ = nullptr;
    unsigned int blocks_cap_ = 0;
    float* d_buf_ = nullptr;
    size_t buf_cap_ = 0;
};

}  // namespace sspsd

using namespace sspsd;

struct sspsd_source {
    Source s;
};

extern "C" {

int32_t sspsd_source_create(int32_t kind, int64_t param, uint64_t seed, int32_t device, void* stream, sspsd_source** out)
{
    if (!out) return SSPSD_EINVAL;
    *out = nullptr;
    sspsd_source* h = new (std::nothrow) sspsd_source;
    if (!h) return SSPSD_ENOMEM;
    int rc = h->s.init(kind, param, seed, device, stream);
    if (rc) {
        delete h;
        return rc;
    }
    *out = h;
    return SSPSD_OK;
}

void sspsd_source_destroy(sspsd_source* h) { delete h; }

int32_t sspsd_source_reset(sspsd_source* h)
{
    if (!h) return SSPSD_EINVAL;
    return h->s.reset();
}

int32_t sspsd_source_generate(sspsd_source* h, float* d_out, size_t n)
{
    if (!h || (n && !d_out)) return SSPSD_EINVAL;
    return h->s.generate(d_out, n, h->s.stream());
}

int32_t sspsd_source_position(const sspsd_source* h, uint64_t* pos)
{
    if (!h || !pos) return SSPSD_EINVAL;
    *pos = h->s.position();
    return SSPSD_OK;
}

int32_t sspsd_source_seek(sspsd_source* h, uint64_t pos)
{
    if (!h) return SSPSD_EINVAL;
    return h->s.seek(pos);
}

int32_t sspsd_cascade_process_source(sspsd_cascade* c, sspsd_source* h, size_t n)
{
    if (!c || !h) return SSPSD_EINVAL;
    if (c->c.device() != h->s.device()) {
        set_error("source and cascade live on different devices");
        return SSPSD_EINVAL;
    }
    // slices bound the scratch buffer; generation and consumption are ordered on the cascade's stream
    const size_t slice = (size_t)1 << 26;
    float* buf = nullptr;
    int rc = h->s.scratch(n < slice ? n : slice, &buf);
    if (rc) return rc;
    while (n) {
        const size_t cnt = n < slice ? n : slice;
        rc = h->s.generate(buf, cnt, c->c.stream());
        if (rc) return rc;
        rc = c->c.process(buf, cnt, SSPSD_MEM_DEVICE);
        if (rc) return rc;
        n -= cnt;
    }
    return SSPSD_OK;
}

}  // extern "C"
