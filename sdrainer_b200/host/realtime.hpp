// realtime.hpp -- measured real-time channel count (BASELINE.json metric, second half: "real-time CW channels per GPU").
//
// The reference's load model is per tick: every stream delivers 93.75 blocks/s (tci/tci.go:157-158 at 48 kS/s / 512,
// scaled to 192 kS/s / 2048), rx.Receiver.run handles one frame at a time, calls Listener.Listen per attached listener
// (rx/receiver.go:388-402) and cw.Decoder.Tick per key state (cw/spectral.go:48-54, cw/decode.go:202).  This harness
// runs that loop for S streams x L listeners end to end against the GPU engine, batch by batch:
//   1. ring copy   every stream's frames of the batch (the []float32 the TCI/Kiwi client hands to Receiver.IQData,
//                  rx/receiver.go:315-334) are copied from ordinary host memory into the C-owned pinned ring
//   2. submit      ONE sdr_submit for all streams (dispatcher form): H2D, K1, K2 with the BoolDebouncer on the device
//   3. collect     sdr_collect of the previous batch: thresholds, noise scalars, packed key bits, peaks
//   4. decode      cw.Decoder.Tick for every (stream, listener, block) key bit on the host cores
// Copy of batch k+1 and decode of batch k-1 overlap the GPU work of batch k (two in-flight slots).  The caller searches
// the largest S whose batch time stays below the batch's signal time (lag < 1 batch).
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "sdrhost.hpp"

#if defined(__AVX2__)
#include <immintrin.h>
#endif

namespace sdrhost {
namespace rt {

// Frame copy into the pinned ring.  The ring is written once and read by the DMA engine only, so the stores bypass the
// cache (non-temporal): no read-for-ownership of the destination lines, a third less host-memory traffic than memcpy --
// host memory bandwidth, shared with the H2D DMA, is what bounds the ring-copy variant.
inline void stream_copy(float *dst, const float *src, size_t n_floats) {
#if defined(__AVX2__)
    if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0 && (n_floats & 31) == 0) {
        const __m256i *s = reinterpret_cast<const __m256i *>(src);
        __m256i *d = reinterpret_cast<__m256i *>(dst);
        const size_t n = n_floats / 8;
        for (size_t i = 0; i < n; i += 4) {
            const __m256i a = _mm256_loadu_si256(s + i), b = _mm256_loadu_si256(s + i + 1), c = _mm256_loadu_si256(s + i + 2),
                          e = _mm256_loadu_si256(s + i + 3);
            _mm256_stream_si256(d + i, a);
            _mm256_stream_si256(d + i + 1, b);
            _mm256_stream_si256(d + i + 2, c);
            _mm256_stream_si256(d + i + 3, e);
        }
        return;
    }
#endif
    std::memcpy(dst, src, n_floats * sizeof(float));
}

class Pool {  // persistent worker threads, parallel_for over contiguous chunks
   public:
    explicit Pool(int n) : n_(n < 1 ? 1 : n) {
        for (int i = 1; i < n_; i++) th_.emplace_back([this, i] { loop(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return n_; }
    void parallel_for(size_t n, const std::function<void(size_t, size_t)> &fn) {
        fn_ = &fn;
        total_ = n;
        {
            std::lock_guard<std::mutex> lk(mu_);
            pending_ = n_ - 1;
            gen_++;
        }
        cv_.notify_all();
        run(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

   private:
    void run(int i) {
        const size_t lo = total_ * (size_t)i / (size_t)n_, hi = total_ * (size_t)(i + 1) / (size_t)n_;
        if (hi > lo) (*fn_)(lo, hi);
    }
    void loop(int i) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            run(i);
            {
                std::lock_guard<std::mutex> lk(mu_);
                pending_--;
            }
            done_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t, size_t)> *fn_ = nullptr;
    size_t total_ = 0;
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct NullWriter : cw::Writer {
    long long chars = 0;
    void Write(const std::string &s) override { chars += (long long)s.size(); }
};

struct Stats {
    double batch_s, copy_s, submit_s, collect_wait_s, decode_s, gpu_ms;
    long long ticks, chars, key_downs;
};

class Harness {
   public:
    // src: n_templates streams x src_blocks blocks x 2N float32 in ordinary host memory; bins: [n_templates][listeners]
    Harness(sdr_engine *e, int fs, int n, int listeners, int s_cap, int blocks_per_batch, int threads, int debounce, const float *src,
            int n_templates, int src_blocks, const int *bins)
        : e_(e), fs_(fs), n_(n), L_(listeners), cap_(s_cap), B_(blocks_per_batch), debounce_(debounce), src_(src), nt_(n_templates),
          sb_(src_blocks), pool_(threads) {
        for (int k = 0; k < 2; k++) {
            void *p = nullptr;
            if (sdr_alloc_pinned(e_, (size_t)cap_ * B_ * 2 * n_ * sizeof(float), &p) != SDR_OK) throw std::runtime_error(sdr_last_error(e_));
            ring_[k] = (float *)p;
        }
        bins_.assign(bins, bins + (size_t)nt_ * L_);
        streams_.resize(cap_);
        for (int s = 0; s < cap_; s++)
            if (sdr_stream_open(e_, fs_, &streams_[s]) != SDR_OK) throw std::runtime_error(sdr_last_error(e_));
        writers_.resize(cap_);
        decoders_.reserve((size_t)cap_ * L_);
        for (int s = 0; s < cap_; s++)
            for (int l = 0; l < L_; l++) decoders_.emplace_back(&writers_[s], fs_, n_);
        works_.resize(cap_);
    }
    ~Harness() {
        for (int k = 0; k < 2; k++) sdr_free_pinned(e_, ring_[k]);
        for (int s : streams_) sdr_stream_close(e_, s);
    }

    // n_batches batches of S streams; the first two fill the pipeline and are not timed
    // ring_copy = false: the frames are produced directly in the pinned ring (a client that receives into C-owned pinned
    // memory, the zero-copy variant INTEGRATION.md describes): the ring is filled by the two untimed batches only
    Stats Run(int S, int n_batches, bool ring_copy = true) {
        using clk = std::chrono::steady_clock;
        auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
        if (S > cap_) S = cap_;
        Stats st{};
        const size_t frame = (size_t)2 * n_;
        sdr_ticket tickets[2] = {0, 0};
        bool inflight[2] = {false, false};
        clk::time_point t_start = clk::now(), t_end = t_start;
        int timed = 0;
        for (int k = 0; k < n_batches + 1; k++) {
            const int slot = k & 1;
            const bool timing = k >= 2 && k < n_batches;  // steady state: a copy, a submit, a collect and a decode per iteration
            if (k == 2) t_start = clk::now();
            if (k == n_batches) t_end = clk::now();
            auto t0 = clk::now();
            if (k < n_batches && (ring_copy || k < 2)) {
                // 1. ring copy (parallel over streams): frames of template s % nt, blocks (k B + b) % src_blocks
                float *ring = ring_[slot];
                pool_.parallel_for((size_t)S, [&](size_t lo, size_t hi) {
                    for (size_t s = lo; s < hi; s++) {
                        const float *tpl = src_ + (size_t)(s % nt_) * sb_ * frame;
                        for (int b = 0; b < B_; b++) {
                            const size_t sblk = ((size_t)k * B_ + b + 7 * s) % (size_t)sb_;
                            stream_copy(ring + ((size_t)s * B_ + b) * frame, tpl + sblk * frame, frame);
                        }
                    }
#if defined(__AVX2__)
                    _mm_sfence();  // the non-temporal stores are globally visible before the submit that follows the join
#endif
                });
            }
            auto t1 = clk::now();
            if (k < n_batches) {
                // 2. one submit for every stream
                for (int s = 0; s < S; s++) {
                    sdr_work &w = works_[s];
                    w = sdr_work{};
                    w.stream = streams_[s];
                    w.n_blocks = B_;
                    w.iq = ring_[slot] + (size_t)s * B_ * frame;
                    w.mem = SDR_MEM_HOST;
                    w.edge_width = rx::defaultEdgeWidth;
                    w.peak_threshold = rx::defaultPeakThreshold;
                    w.n_listeners = L_;
                    w.listener_bins = bins_.data() + (size_t)(s % nt_) * L_;
                    w.signal_debounce = debounce_;
                }
                if (sdr_submit(e_, works_.data(), S, SDR_NO_TAPS | SDR_NO_RAW_KEYS, &tickets[slot]) != SDR_OK)
                    throw std::runtime_error(sdr_last_error(e_));
                inflight[slot] = true;
            }
            auto t2 = clk::now();
            // 3. collect the previous batch, 4. decode it
            const int prev = slot ^ 1;
            double wait_s = 0, dec_s = 0;
            if (k >= 1 && inflight[prev]) {
                sdr_result r;
                if (sdr_collect(e_, tickets[prev], 1, &r) != SDR_OK) throw std::runtime_error(sdr_last_error(e_));
                auto t3 = clk::now();
                wait_s = secs(t2, t3);
                std::atomic<long long> downs{0};
                pool_.parallel_for((size_t)S, [&](size_t lo, size_t hi) {
                    long long d = 0;
                    for (size_t s = lo; s < hi; s++) {
                        const int b0 = r.work_block_offset[s];
                        cw::Decoder *dec = &decoders_[s * (size_t)L_];
                        for (int b = 0; b < B_; b++) {
                            const uint32_t *bits = r.key_bits + (size_t)(b0 + b) * r.key_words;
                            for (int l = 0; l < L_; l++) {
                                const bool key = ((bits[l >> 5] >> (l & 31)) & 1u) != 0;
                                d += key;
                                dec[l].Tick(key);  // cw/decode.go:202
                            }
                        }
                    }
                    downs += d;
                });
                auto t4 = clk::now();
                dec_s = secs(t3, t4);
                if (timing) {
                    st.gpu_ms += r.gpu_ms;
                    st.key_downs += downs.load();
                    st.ticks += (long long)S * B_ * L_;
                }
                sdr_release(e_, tickets[prev]);
                inflight[prev] = false;
            }
            if (timing) {
                st.copy_s += secs(t0, t1);
                st.submit_s += secs(t1, t2);
                st.collect_wait_s += wait_s;
                st.decode_s += dec_s;
                timed++;
            }
        }
        const double total = secs(t_start, t_end);
        const int nb = timed > 0 ? timed : 1;
        st.batch_s = total / nb;
        st.copy_s /= nb;
        st.submit_s /= nb;
        st.collect_wait_s /= nb;
        st.decode_s /= nb;
        st.gpu_ms /= nb;
        for (auto &w : writers_) st.chars += w.chars;
        return st;
    }

   private:
    sdr_engine *e_;
    int fs_, n_, L_, cap_, B_, debounce_;
    const float *src_;
    int nt_, sb_;
    Pool pool_;
    float *ring_[2] = {nullptr, nullptr};
    std::vector<int> bins_, streams_;
    std::vector<NullWriter> writers_;
    std::vector<cw::Decoder> decoders_;
    std::vector<sdr_work> works_;
};

}  // namespace rt
}  // namespace sdrhost
