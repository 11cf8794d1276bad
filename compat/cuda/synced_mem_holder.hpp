// compat/cuda/synced_mem_holder.hpp -- header-compatible replacement of the reference's
// include/cuda/synced_mem_holder.hpp:10-65 (orb_cuda::SyncedMem<T>: pinned host buffer + device mirror + own stream).
//
// Same methods and the same PUBLIC data members (count_, capacity_, cpu_data_, gpu_data_, pitch_, cu_stream_,
// cu_error_) that Frame.cpp / Tracking.cpp / ORBmatcher.cpp touch directly (src/Frame.cpp:126, src/cuda/orb_gpu.cpp:85).
// Differences, both deliberate: (1) header-only; (2) storage is reference-counted, so the implicit copies the SLAM core
// makes (`mCurrentFrame = Frame(...)`, src/Tracking.cpp:292) no longer double-free.
// A SyncedMem can also be a non-owning VIEW of device memory owned by a jsfe handle (ORB_GPU::image_).
#ifndef JSFE_COMPAT_SYNCED_MEM_HOLDER_HPP
#define JSFE_COMPAT_SYNCED_MEM_HOLDER_HPP

#include <cuda_runtime_api.h>

#include <cstring>
#include <memory>

namespace orb_cuda {

template <typename Dtype>
class SyncedMem {
    // The stream and the buffers are reference-counted SEPARATELY: copies of a SyncedMem share both, and a resize() of one copy
    // replaces only its buffers -- the stream stays alive for every copy that still holds the handle (cu_stream_ is public).
    struct Stream {
        cudaStream_t s = nullptr;
        Stream() { cudaStreamCreate(&s); }
        ~Stream() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
    };
    struct Block {
        Dtype* cpu = nullptr;
        Dtype* gpu = nullptr;
        std::shared_ptr<Stream> stream;     // work on this stream may still use the buffers: drain it before they go
        ~Block() {
            if (stream && stream->s) cudaStreamSynchronize(stream->s);
            if (cpu) cudaFreeHost(cpu);
            if (gpu) cudaFree(gpu);
        }
    };
    std::shared_ptr<Stream> stream_;
    std::shared_ptr<Block> block_;

public:
    SyncedMem(void) : stream_(std::make_shared<Stream>()), block_(std::make_shared<Block>()) {
        count_ = 0; capacity_ = 0; cpu_data_ = nullptr; gpu_data_ = nullptr; pitch_ = 0; cu_error_ = cudaSuccess;
        block_->stream = stream_;
        cu_stream_ = stream_->s;
    }
    // non-owning device view (used for ORB_GPU::image_)
    static SyncedMem view(Dtype* dev, int count, size_t pitch) {
        SyncedMem m;
        m.gpu_data_ = dev; m.count_ = count; m.capacity_ = 0; m.pitch_ = pitch;
        return m;
    }
    ~SyncedMem() = default;

    void resize(int count) {
        count_ = count;
        if (capacity_ < count_) {
            auto nb = std::make_shared<Block>();
            nb->stream = stream_;                                     // same stream, new buffers (the old block drains it when it dies)
            cu_error_ = cudaMallocHost((void**)&nb->cpu, sizeof(Dtype) * count_);
            if (cu_error_ == cudaSuccess) cu_error_ = cudaMalloc((void**)&nb->gpu, sizeof(Dtype) * count_);
            capacity_ = count_;
            block_ = nb;
            cpu_data_ = nb->cpu; gpu_data_ = nb->gpu; cu_stream_ = stream_->s;
        }
    }
    void resize_pitched(size_t width, size_t height) {
        resize((int)(width * height));
        pitch_ = width * sizeof(Dtype);
    }
    Dtype* cpu_data() { return cpu_data_; }
    Dtype* gpu_data() { return gpu_data_; }

    void to_cpu(void) { to_cpu(count_); }
    void to_gpu(void) { to_gpu(count_); }
    void to_cpu(int count) { cu_error_ = cudaMemcpy(cpu_data_, gpu_data_, sizeof(Dtype) * count, cudaMemcpyDeviceToHost); }
    void to_gpu(int count) { cu_error_ = cudaMemcpy(gpu_data_, cpu_data_, sizeof(Dtype) * count, cudaMemcpyHostToDevice); }
    void to_cpu_async(void) { to_cpu_async(cu_stream_, count_); }
    void to_gpu_async(void) { to_gpu_async(cu_stream_, count_); }
    void to_cpu_async(cudaStream_t& s) { to_cpu_async(s, count_); }
    void to_gpu_async(cudaStream_t& s) { to_gpu_async(s, count_); }
    void to_cpu_async(int count) { to_cpu_async(cu_stream_, count); }
    void to_gpu_async(int count) { to_gpu_async(cu_stream_, count); }
    void to_cpu_async(cudaStream_t& s, int count) { cu_error_ = cudaMemcpyAsync(cpu_data_, gpu_data_, sizeof(Dtype) * count, cudaMemcpyDeviceToHost, s); }
    void to_gpu_async(cudaStream_t& s, int count) { cu_error_ = cudaMemcpyAsync(gpu_data_, cpu_data_, sizeof(Dtype) * count, cudaMemcpyHostToDevice, s); }
    void sync_stream(void) { cu_error_ = cudaStreamSynchronize(cu_stream_); }
    void set_zero_gpu(void) { if (gpu_data_) cu_error_ = cudaMemset(gpu_data_, 0, sizeof(Dtype) * count_); }
    void set_zero_gpu_async(void) { if (gpu_data_) cu_error_ = cudaMemsetAsync(gpu_data_, 0, sizeof(Dtype) * count_, cu_stream_); }
    void set_zero_cpu(void) { if (cpu_data_) memset(cpu_data_, 0, sizeof(Dtype) * count_); }

    // public, as in the reference
    int count_;
    int capacity_;
    Dtype* cpu_data_;
    Dtype* gpu_data_;
    size_t pitch_;
    cudaStream_t cu_stream_;
    cudaError_t cu_error_;
};

}  // namespace orb_cuda
#endif
