// sdrhost.hpp -- host side above the C ABI: a C++ mirror of the reference's Go interface for the hot path
// (packages dsp, cw, rx), same names, argument meaning and error behaviour, so that the parity tests read like
// the reference's own tests.  The Go toolchain is absent from the build image, hence C++.
//
// What runs where: every per-sample / per-bin loop is on the GPU behind include/sdrgpu.h; what stays here is what
// the reference keeps sequential and stateful -- BoolDebouncer, cw.Decoder, PeaksTable, ListenerPool, Listener
// timeouts and the frame bookkeeping of rx.Receiver.run.  Nothing in this file computes a spectrum on the CPU.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sdrgpu.h"

namespace sdrhost {

// ============================================================ dsp =====================================
namespace dsp {

constexpr double BinFrom = -0.5, BinCenter = 0.0, BinTo = 0.5;  // dsp/fft.go:89-93

inline int64_t go_int(double x) {  // Go int(float64) on amd64
    if (std::isnan(x) || x >= 9223372036854775808.0 || x < -9223372036854775808.0) return INT64_MIN;
    return (int64_t)x;
}

struct Peak {  // dsp.Peak[float32,int], dsp/fft.go:179-213
    int From = 0, To = 0;
    int64_t FromFrequency = 0, ToFrequency = 0, SignalFrequency = 0;
    float SignalValue = 0;
    int SignalBin = 0;
    int Center() const { return From + ((To - From) / 2); }
    int64_t WidthHz() const { return ToFrequency - FromFrequency; }
    int64_t CenterFrequency() const { return FromFrequency + (WidthHz() / 2); }
    int Width() const { return (To - From) + 1; }
    bool ContainsBin(int bin) const { return From >= bin && To <= bin; }  // sic, dsp/fft.go:211-213
};

class FrequencyMapping {  // dsp/fft.go:95-135
   public:
    FrequencyMapping(int sampleRate, int blockSize, int64_t centerFrequency)
        : sampleRate_(sampleRate), blockSize_(blockSize), binSize_((double)sampleRate / (double)blockSize) {
        SetCenterFrequency(centerFrequency);
    }
    void SetCenterFrequency(int64_t f) {
        centerFrequency_ = f;
        fromFrequency_ = f - sampleRate_ / 2;
    }
    int64_t BinToFrequency(int bin, double location) const {
        const double locationDelta = binSize_ * location;
        return (int64_t)((uint64_t)fromFrequency_ + (uint64_t)go_int((double)bin * binSize_ + locationDelta));
    }
    int FrequencyToBin(int64_t frequency) const {
        int64_t bin = go_int(((double)frequency - (double)fromFrequency_) / binSize_);
        return (int)std::max<int64_t>(0, std::min<int64_t>(bin, blockSize_ - 1));
    }

   private:
    int sampleRate_, blockSize_;
    double binSize_;
    int64_t centerFrequency_ = 0, fromFrequency_ = 0;
};

// dsp.PeakCenterCorrection (dsp/fft.go:292-309) from the three cumulation values the GPU returns
inline double PeakCenterCorrection(int bin, int blockSize, float y1f, float y2f, float y3f) {
    if (bin <= 0 || bin >= blockSize - 1) return 0;
    const double y1 = std::fabs((double)y1f), y2 = std::fabs((double)y2f), y3 = std::fabs((double)y3f);
    return (y3 - y1) / (2 * (2 * y2 - y1 - y3));
}

// dsp.Peak from the GPU's record: the integer / float64 frequency maths of dsp/fft.go:264-268
inline Peak FromGpuPeak(const sdr_peak &g, int blockSize, const FrequencyMapping &m) {
    Peak p;
    p.From = g.from;
    p.To = g.to;
    p.SignalBin = g.signal_bin;
    p.SignalValue = g.signal_value;
    p.FromFrequency = m.BinToFrequency(p.From, BinFrom);
    p.ToFrequency = m.BinToFrequency(p.To, BinTo);
    p.SignalFrequency = m.BinToFrequency(p.SignalBin, PeakCenterCorrection(p.SignalBin, blockSize, g.y1, g.y2, g.y3));
    return p;
}

class BoolDebouncer {  // dsp/dsp.go:139-182
   public:
    explicit BoolDebouncer(int threshold) : threshold_(threshold) {}
    void SetThreshold(int t) { threshold_ = t; }
    int Threshold() const { return threshold_; }
    bool Debounce(bool rawState) {
        if (threshold_ < 2) return rawState;
        if (rawState != lastRawState_) stateCount_ = 1;
        else stateCount_++;
        lastRawState_ = rawState;
        if (stateCount_ >= threshold_) {
            if (rawState != effectiveState_) effectiveState_ = rawState;
        }
        return effectiveState_;
    }

   private:
    int threshold_;
    bool effectiveState_ = false, lastRawState_ = false;
    int stateCount_ = 0;
};

// dsp.FFT / FindNoiseFloor / FindPeaks with the reference's signatures, executed by the engine.
class FFT {
   public:
    explicit FFT(sdr_engine *e) : e_(e) {}
    // dsp/fft.go:23 with the receiver's shiftedMagnitude projection; spectrum/psd must have blockSize entries
    void IQToSpectrumAndPSD(std::vector<float> &spectrum, std::vector<float> &psd, const std::vector<float> &iq, int blockSize) {
        if ((int)spectrum.size() != blockSize)  // the reference panics here (dsp/fft.go:28-30)
            throw std::invalid_argument("the spectrum slice must have the same length as the FFT's result");
        if (sdr_dsp_iq_to_spectrum_and_psd(e_, iq.data(), 1, spectrum.data(), psd.data()) != SDR_OK)
            throw std::runtime_error(sdr_last_error(e_));
    }

   private:
    sdr_engine *e_;
};

}  // namespace dsp

// ============================================================ cw ======================================
namespace cw {

// Morse table: ITU-R M.1677-1 plus the entries the reference's tests pin ('a', '/', U+00A7 = eight dits,
// and 'ä' through the golden strings, cw/decode_test.go:26-28,184-192).  The reference takes it from
// github.com/ftl/digimodes (not vendored).
inline const std::map<std::string, char32_t> &DecodeTable() {
    static const std::map<std::string, char32_t> t = {
        {".-", U'a'},     {"-...", U'b'},   {"-.-.", U'c'},   {"-..", U'd'},    {".", U'e'},      {"..-.", U'f'},
        {"--.", U'g'},    {"....", U'h'},   {"..", U'i'},     {".---", U'j'},   {"-.-", U'k'},    {".-..", U'l'},
        {"--", U'm'},     {"-.", U'n'},     {"---", U'o'},    {".--.", U'p'},   {"--.-", U'q'},   {".-.", U'r'},
        {"...", U's'},    {"-", U't'},      {"..-", U'u'},    {"...-", U'v'},   {".--", U'w'},    {"-..-", U'x'},
        {"-.--", U'y'},   {"--..", U'z'},   {"-----", U'0'},  {".----", U'1'},  {"..---", U'2'},  {"...--", U'3'},
        {"....-", U'4'},  {".....", U'5'},  {"-....", U'6'},  {"--...", U'7'},  {"---..", U'8'},  {"----.", U'9'},
        {".-.-.-", U'.'}, {"--..--", U','}, {"..--..", U'?'}, {"-..-.", U'/'},  {"-...-", U'='},  {".-.-.", U'+'},
        {"-....-", U'-'}, {".--.-.", U'@'}, {"---...", U':'}, {"-.-.-.", U';'}, {".----.", U'\''}, {".-..-.", U'"'},
        {"-.--.", U'('},  {"-.--.-", U')'}, {"-.-.--", U'!'}, {".-...", U'&'},  {"..--.-", U'_'}, {"...-..-", U'$'},
        {".-.-", U'ä'}, {"---.", U'ö'}, {"..--", U'ü'}, {"........", U'§'},
    };
    return t;
}

inline void AppendUtf8(std::string &s, char32_t r) {
    if (r < 0x80) s.push_back((char)r);
    else if (r < 0x800) {
        s.push_back((char)(0xC0 | (r >> 6)));
        s.push_back((char)(0x80 | (r & 0x3F)));
    } else {
        s.push_back((char)(0xE0 | (r >> 12)));
        s.push_back((char)(0x80 | ((r >> 6) & 0x3F)));
        s.push_back((char)(0x80 | (r & 0x3F)));
    }
}

struct Writer {  // io.Writer
    virtual ~Writer() = default;
    virtual void Write(const std::string &utf8) = 0;
};

class AdaptiveThreshold {  // cw/decode.go:360-431
   public:
    explicit AdaptiveThreshold(double preset = 1) : preset_(preset) { Reset(); }
    void Reset() {
        low_ = preset_;
        high_ = 3 * low_;
        last_ = low_;
        update();
    }
    void Preset(double p) {
        preset_ = p;
        Reset();
    }
    void Put(double duration) {
        const double highFactor = 2, avgWeight = 0.75, currentWeight = 1.0 - 0.75;
        if (duration >= low_ * upperBound_) return;
        if (last_ >= duration * highFactor) {
            low_ = avgWeight * low_ + currentWeight * duration;
            high_ = avgWeight * high_ + currentWeight * last_;
        } else if (duration >= last_ * highFactor) {
            low_ = avgWeight * low_ + currentWeight * last_;
            high_ = avgWeight * high_ + currentWeight * duration;
        }
        last_ = duration;
        update();
    }
    double Get() const { return threshold_; }
    double Low() const { return low_; }
    double High() const { return high_; }

   private:
    void update() { threshold_ = std::sqrt(low_ * high_); }
    double preset_, upperBound_ = 10, low_ = 0, high_ = 0, last_ = 0, threshold_ = 0;
};

class Decoder {  // cw/decode.go:108-358
   public:
    static constexpr int maxSymbolCount = 8;
    static constexpr char32_t unknownCharacter = 0xA6;
    Decoder(Writer *out, int sampleRate, int blockSize)
        : out_(out), tickSeconds_((double)blockSize / (double)sampleRate), wpm_(20), abortDecodeAfterDits_(10) {
        const double ditTime = wpmToDit(wpm_);
        onThreshold_ = AdaptiveThreshold(ditTime);
        offThreshold_ = AdaptiveThreshold(ditTime);
    }
    void Reset() {  // :166-170 -- lastState and currentCharInvalid survive, as in the reference
        presetWPM(20);
        Clear();
        onThreshold_.Reset();
    }
    void Clear() {
        decoding_ = false;
        currentChar_.clear();
        ticks_ = 0;
        onStart_ = 0;
        offStart_ = 0;
    }
    void Tick(bool state) {  // :202-250
        ticks_++;
        const double now = ticks_;
        if (state != lastState_) {
            if (state) {
                onStart_ = now;
                onRisingEdge(now - offStart_);
            } else {
                offStart_ = now;
                onFallingEdge(now - onStart_);
            }
            decoding_ = true;
        }
        lastState_ = state;
        const double currentDuration = state ? now - onStart_ : now - offStart_;
        const double upperBound = offThreshold_.Get() * (double)abortDecodeAfterDits_;
        if (decoding_ && currentDuration > upperBound) {
            decoding_ = false;
            decodeCurrentChar();
        }
    }
    void stop() { decodeCurrentChar(); }  // :356-358
    double WPM() const { return wpm_; }

   private:
    double wpmToDit(double wpm) const { return std::ceil((60.0 / (50.0 * wpm)) / tickSeconds_); }
    double ditToWPM(double ditTicks) const { return 60.0 / (50.0 * (ditTicks * tickSeconds_)); }
    void presetWPM(int wpm) {
        wpm_ = (double)wpm;
        const double ditTime = wpmToDit(wpm_);
        onThreshold_.Preset(ditTime);
        offThreshold_.Preset(ditTime);
    }
    void onRisingEdge(double offDuration) {  // :252-275
        if (offDuration < 2.0) return;
        offThreshold_.Put(offDuration);
        const double threshold = offThreshold_.Get();
        const double upperThreshold = 4.5 * offThreshold_.Low();
        if (offDuration >= upperThreshold) {
            decodeCurrentChar();
            writeToOutput(U' ');
        } else if (offDuration >= threshold) {
            decodeCurrentChar();
        }
    }
    void onFallingEdge(double onDuration) {  // :277-297
        if (onDuration < 2.0) return;
        onThreshold_.Put(onDuration);
        const double threshold = onThreshold_.Get();
        const double upperThreshold = 2 * onThreshold_.High();
        if (onDuration >= upperThreshold) {
            currentCharInvalid_ = true;
        } else if (onDuration >= threshold) {
            appendSymbol('-');
            wpm_ = (wpm_ + ditToWPM(onThreshold_.Low())) / 2.0;
        } else {
            appendSymbol('.');
        }
    }
    void appendSymbol(char s) {  // :306-312
        if ((int)currentChar_.size() >= maxSymbolCount) decodeCurrentChar();
        currentChar_.push_back(s);
    }
    void decodeCurrentChar() {  // :314-349
        if (currentChar_.empty()) return;
        if (currentCharInvalid_) {
            currentCharInvalid_ = false;
            currentChar_.clear();
            writeToOutput(unknownCharacter);
            return;
        }
        auto it = DecodeTable().find(currentChar_);
        writeToOutput(it != DecodeTable().end() ? it->second : unknownCharacter);
        currentChar_.clear();
    }
    void writeToOutput(char32_t r) {
        if (!out_) return;
        std::string s;
        AppendUtf8(s, r);
        out_->Write(s);
    }

    Writer *out_;
    double tickSeconds_, ticks_ = 0;
    bool lastState_ = false;
    double onStart_ = 0, offStart_ = 0, wpm_;
    bool decoding_ = false;
    int abortDecodeAfterDits_;
    std::string currentChar_;
    bool currentCharInvalid_ = false;
    AdaptiveThreshold onThreshold_, offThreshold_;
};

class SpectralDemodulator {  // cw/spectral.go
   public:
    SpectralDemodulator(Writer *out, int sampleRate, int blockSize) : signalDebouncer_(1), decoder_(out, sampleRate, blockSize) {}
    void SetSignalDebounce(int d) { signalDebouncer_.SetThreshold(d); }
    void Reset() { decoder_.Reset(); }
    // cw/spectral.go:48-54; `state` may come precomputed from the GPU (sdr_result.keys)
    bool Tick(float value, float threshold) { return TickState(value > threshold); }
    bool TickState(bool state) {
        const bool debounced = signalDebouncer_.Debounce(state);
        decoder_.Tick(debounced);
        return debounced;
    }
    // the debouncer already ran on the device (sdr_result.key_bits): only cw.Decoder.Tick is left for the host
    void TickDebounced(bool debounced) { decoder_.Tick(debounced); }

   private:
    dsp::BoolDebouncer signalDebouncer_;
    Decoder decoder_;
};

// cw.AudioDemodulator (cw/audio.go): the per-sample channel of the reference becomes a buffered Write; whole
// Goertzel blocks go to the GPU bank in one call, debouncer and decoder stay here.
class AudioDemodulator {
   public:
    AudioDemodulator(Writer *out, double pitch, int sampleRate, int maxBlocks = 4096) : debouncer_(3), sampleRate_(sampleRate) {
        sdr_goertzel_config c{};
        c.device = 0;
        c.sample_rate = sampleRate;
        c.n_filters = 1;
        c.pitch = &pitch;
        c.blocksize_ratio = 0.005;
        c.max_blocks = maxBlocks;
        if (sdr_goertzel_create(&c, &bank_) != SDR_OK) throw std::runtime_error(sdr_goertzel_last_error(nullptr));
        blocksize_ = sdr_goertzel_blocksize(bank_, 0);
        maxBlocks_ = maxBlocks;
        decoder_.reset(new Decoder(out, sampleRate, blocksize_));
    }
    ~AudioDemodulator() { sdr_goertzel_destroy(bank_); }
    int Blocksize() const { return blocksize_; }
    void SetScale(double s) { scale_ = (float)s; }
    void SetMaxScale(double s) { maxScale_ = s; }
    void SetChannelCount(int c) { channelCount_ = c; }
    void SetDebounceThreshold(int t) { debouncer_.SetThreshold(t); }
    int Write(const float *buf, int n) {  // cw/audio.go:149-158 + run() :169-204
        for (int i = 0; i < n; i++)
            if ((i % channelCount_) == 0) pending_.push_back(buf[i]);
        int nb = (int)(pending_.size() / blocksize_);
        while (nb > 0) {
            const int take = std::min(nb, maxBlocks_);
            std::vector<double> mag(take);
            std::vector<uint8_t> st(take);
            const float *ptrs[1] = {pending_.data()};
            if (sdr_goertzel_process_audio(bank_, ptrs, &take, &scale_, maxScale_, mag.data(), st.data(), take) != SDR_OK)
                throw std::runtime_error(sdr_goertzel_last_error(bank_));
            for (int b = 0; b < take; b++) {
                lastMagnitude_ = mag[b];
                decoder_->Tick(debouncer_.Debounce(st[b] != 0));
            }
            pending_.erase(pending_.begin(), pending_.begin() + (size_t)take * blocksize_);
            nb -= take;
        }
        return n;
    }
    void Close() { decoder_->stop(); }
    double LastMagnitude() const { return lastMagnitude_; }

   private:
    sdr_goertzel_bank *bank_ = nullptr;
    dsp::BoolDebouncer debouncer_;
    std::unique_ptr<Decoder> decoder_;
    std::vector<float> pending_;
    int sampleRate_, blocksize_ = 0, maxBlocks_ = 0, channelCount_ = 1;
    float scale_ = 1;
    double maxScale_ = 12, lastMagnitude_ = 0;
};

}  // namespace cw

// ============================================================ rx ======================================
namespace rx {

constexpr int iqBufferSize = 100, cumulationSize = 100, peakPadding = 0;  // rx/receiver.go:15-27
constexpr float defaultPeakThreshold = 15;
constexpr int defaultEdgeWidth = 70, defaultListenerPoolSize = 30;
constexpr int64_t kSecond = 1000000000ll;
constexpr int64_t defaultPeakTimeout = 120 * kSecond;        // rx/peaks.go:11
constexpr int64_t defaultSilenceTimeout = 20 * kSecond;      // rx/listener.go:15
constexpr int64_t defaultAttachmentTimeout = 120 * kSecond;  // rx/listener.go:16

struct Clock {  // rx/receiver.go:29-31
    virtual ~Clock() = default;
    virtual int64_t Now() const = 0;  // nanoseconds
};
struct ManualClock : Clock {  // rx/receiver.go:41-55
    int64_t now = 0;
    int64_t Now() const override { return now; }
    void Set(int64_t t) { now = t; }
    void Add(int64_t d) { now += d; }
};

enum class ReceiverMode { Decode, Strain };

struct Reporter {  // rx/rx.go:11-17 (the two events that originate on this path)
    virtual ~Reporter() = default;
    virtual void ListenerActivated(const std::string &listener, int64_t frequency) = 0;
    virtual void ListenerDeactivated(const std::string &listener, int64_t frequency) = 0;
};

class PeaksTable {  // rx/peaks.go
   public:
    enum State { peakNone = 0, peakNew, peakActive, peakInactive };
    struct Entry {
        dsp::Peak *Peak;
        State state;
        int64_t since;
    };
    PeaksTable(int size, const Clock *clock) : bins_(size, nullptr), clock_(clock) {}
    void ForcePut(dsp::Peak *p) { put(p, true); }
    bool Put(dsp::Peak *p) { return put(p, false); }  // false: refused (overlaps an active / inactive peak)
    dsp::Peak *Get(int bin) const {
        if (bin < 0 || bin >= (int)bins_.size() || !bins_[bin]) return nullptr;
        return bins_[bin]->Peak;
    }
    const Entry *Bin(int bin) const { return bins_[bin]; }
    void Cleanup() {  // :127-147
        const int64_t now = clock_->Now();
        size_t i = 0;
        while (i < bins_.size()) {
            Entry *p = bins_[i];
            i++;
            if (!p) continue;
            if (p->state == peakActive) continue;
            if (now - p->since < peakTimeout) continue;
            const int to = p->Peak->To;
            clear(p->Peak->From, to);
            i = (size_t)to + 1;
        }
    }
    void Reset() {
        std::fill(bins_.begin(), bins_.end(), nullptr);
        for (auto &en : entries_)
            if (onDrop) onDrop(en->Peak);
        entries_.clear();
    }
    // called with the Peak of every entry that left the table (the Go code leaves this to the garbage collector)
    std::function<void(dsp::Peak *)> onDrop;
    void Activate(dsp::Peak *p) {
        Entry *e = getInternal(p);
        if (!e) return;  // the reference would nil-deref
        if (e->state != peakNew && e->state != peakInactive) return;
        e->state = peakActive;
    }
    void Deactivate(dsp::Peak *p) {
        Entry *e = getInternal(p);
        if (!e) return;
        if (e->state != peakActive) return;
        e->state = peakInactive;
    }
    // rx/peaks.go:183-207.  The reference probes len/2 random bins first (unseeded math/rand); here the probe
    // uses a seeded xorshift64* shared with the oracle, or is skipped (deterministic: lowest new peak).
    dsp::Peak *FindNext() {
        if (!deterministic) {
            for (size_t i = 0; i < bins_.size() / 2; i++) {
                Entry *p = bins_[(size_t)(next() % (uint64_t)bins_.size())];
                if (!p || p->state != peakNew) continue;
                return p->Peak;
            }
        }
        for (Entry *p : bins_) {
            if (!p || p->state != peakNew) continue;
            return p->Peak;
        }
        return nullptr;
    }
    int64_t peakTimeout = defaultPeakTimeout;
    bool deterministic = true;
    uint64_t rng = 1;

   private:
    bool put(dsp::Peak *p, bool force) {  // :46-100
        int clearFrom = -1, clearTo = -1;
        for (int i = std::max(0, p->From); i <= std::min(p->To, (int)bins_.size() - 1); i++) {
            Entry *e = bins_[i];
            if (!e) continue;
            if (!force && (e->state == peakActive || e->state == peakInactive)) return false;
            if (clearFrom == -1) clearFrom = e->Peak->From;
            clearTo = e->Peak->To;
        }
        if (clearFrom > -1 && clearTo > -1) clear(clearFrom, clearTo);
        entries_.emplace_back(new Entry{p, peakNew, clock_->Now()});
        Entry *ne = entries_.back().get();
        for (int i = std::max(0, p->From); i <= std::min(p->To, (int)bins_.size() - 1); i++) bins_[i] = ne;
        return true;
    }
    void clear(int from, int to) {
        std::vector<Entry *> gone;
        for (int i = std::max(0, from); i <= std::min(to, (int)bins_.size() - 1); i++) {
            Entry *en = bins_[i];
            bins_[i] = nullptr;
            if (en && (gone.empty() || gone.back() != en)) gone.push_back(en);
        }
        for (Entry *en : gone) dropIfUnreferenced(en);
    }
    void dropIfUnreferenced(Entry *en) {
        // a replaced peak can be wider than the range being cleared (rx/peaks.go clears exactly [from, to]): the entry
        // stays alive while any bin still points at it
        for (int i = std::max(0, en->Peak->From); i <= std::min(en->Peak->To, (int)bins_.size() - 1); i++)
            if (bins_[i] == en) return;
        for (size_t k = 0; k < entries_.size(); k++)
            if (entries_[k].get() == en) {
                if (onDrop) onDrop(en->Peak);
                entries_[k].swap(entries_.back());
                entries_.pop_back();
                return;
            }
    }
    Entry *getInternal(dsp::Peak *p) {
        Entry *e = bins_[p->From];
        if (!e) return nullptr;
        if (e->Peak->To != p->To) return nullptr;
        return e;
    }
    uint64_t next() {
        uint64_t x = rng;
        x ^= x >> 12;
        x ^= x << 25;
        x ^= x >> 27;
        rng = x;
        return x * 0x2545F4914F6CDD1Dull;
    }
    std::vector<Entry *> bins_;
    std::vector<std::unique_ptr<Entry>> entries_;
    const Clock *clock_;
};

class IDPool {  // rx/listener.go:149-177
   public:
    IDPool(int size, const std::string &prefix) {
        for (int i = 0; i < size; i++) ids_.push_back(prefix + std::to_string(size - i));
    }
    void Push(const std::string &id) { ids_.push_back(id); }
    bool Pop(std::string &id) {
        if (ids_.empty()) return false;
        id = ids_.back();
        ids_.pop_back();
        return true;
    }
    size_t Len() const { return ids_.size(); }

   private:
    std::vector<std::string> ids_;
};

// TextProcessor reduced to what the hot path touches: the sink of decoded text and lastWrite for the silence
// timeout (rx/text_processor.go:151-218).  Callsign extraction is out of scope.
class TextProcessor : public cw::Writer {
   public:
    explicit TextProcessor(const Clock *clock) : clock_(clock), lastWrite_(clock->Now()) {}
    void Restart() { lastWrite_ = clock_->Now(); }
    int64_t LastWrite() const { return lastWrite_; }
    void Write(const std::string &s) override {
        lastWrite_ = clock_->Now();
        text_ += s;
    }
    // rx/text_processor.go:194-199: after defaultWriteTimeout (5 s) of silence the reference flushes its callsign
    // window; the callsign logic itself is out of scope, the tick that triggers it is mirrored (and counted)
    void CheckWriteTimeout() {
        if (clock_->Now() - lastWrite_ > 5 * 1000000000ll) writeTimeouts++;
    }
    int writeTimeouts = 0;
    const std::string &Text() const { return text_; }

   private:
    const Clock *clock_;
    int64_t lastWrite_;
    std::string text_;
};

class Listener {  // rx/listener.go:19-147
   public:
    Listener(const std::string &id, const Clock *clock, Reporter *reporter, int sampleRate, int blockSize)
        : id_(id), clock_(clock), reporter_(reporter), textProcessor_(clock), demodulator_(&textProcessor_, sampleRate, blockSize) {}
    const std::string &ID() const { return id_; }
    void SetSilenceTimeout(int64_t t) { silenceTimeout_ = t; }
    void SetAttachmentTimeout(int64_t t) { attachmentTimeout_ = t; }
    void SetSignalDebounce(int d) { demodulator_.SetSignalDebounce(d); }
    void Attach(dsp::Peak *peak) {
        peak_ = peak;
        lastAttach_ = clock_->Now();
        demodulator_.Reset();
        textProcessor_.Restart();
        if (reporter_) reporter_->ListenerActivated(id_, peak_->SignalFrequency);
    }
    bool Attached() const { return peak_ != nullptr; }
    void Detach() {
        const int64_t f = peak_->SignalFrequency;
        peak_ = nullptr;
        if (reporter_) reporter_->ListenerDeactivated(id_, f);
    }
    dsp::Peak *Peak() const { return peak_; }
    int SignalBin() const { return Attached() ? peak_->SignalBin : 0; }
    bool TimeoutExceeded() const {
        const int64_t now = clock_->Now();
        return (now - lastAttach_ > attachmentTimeout_) || (now - textProcessor_.LastWrite() > silenceTimeout_);
    }
    // rx/listener.go:142-147 with the key state already evaluated on the GPU (value > threshold)
    bool ListenState(bool state) {
        if (!Attached()) return false;
        const bool k = demodulator_.TickState(state);
        if (recordKeys) keys_.push_back(k ? 1 : 0);
        return k;
    }
    bool ListenDebounced(bool debounced) {  // same, for key states debounced on the device
        if (!Attached()) return false;
        demodulator_.TickDebounced(debounced);
        if (recordKeys) keys_.push_back(debounced ? 1 : 0);
        return debounced;
    }
    void CheckWriteTimeout() { textProcessor_.CheckWriteTimeout(); }  // rx/listener.go:138-140
    int slot = -1;      // stable position of this listener in the device's per-stream listener table (pool slot)
    bool fresh = true;  // bound since the last submit: the device debouncer of its slot starts from zero
    const std::string &Text() const { return textProcessor_.Text(); }
    const std::vector<uint8_t> &Keys() const { return keys_; }
    bool recordKeys = true;  // test hook: keep the debounced key stream (grows with the run time)
    int tapIndex = -1;  // column of this listener in the batch being consumed

   private:
    std::string id_;
    const Clock *clock_;
    Reporter *reporter_;
    TextProcessor textProcessor_;
    cw::SpectralDemodulator demodulator_;
    dsp::Peak *peak_ = nullptr;
    int64_t lastAttach_ = 0, silenceTimeout_ = defaultSilenceTimeout, attachmentTimeout_ = defaultAttachmentTimeout;
    std::vector<uint8_t> keys_;
};

class ListenerPool {  // rx/listener.go:179-270
   public:
    using Factory = std::function<Listener *(const std::string &)>;
    ListenerPool(int size, const std::string &prefix, Factory f) : size_(size), ids_(size, prefix), factory_(std::move(f)) {}
    int Size() const { return size_; }
    bool Available() const { return (int)listeners_.size() < size_; }
    void Reset() {
        for (Listener *l : listeners_) {
            if (l->Attached()) l->Detach();
            ids_.Push(l->ID());
        }
        listeners_.clear();
    }
    Listener *BindNext() {
        if ((int)listeners_.size() == size_) return nullptr;
        std::string id;
        if (!ids_.Pop(id)) return nullptr;
        Listener *l = factory_(id);
        listeners_.push_back(l);
        return l;
    }
    void Release(Listener *listener) {
        int index = -1;
        for (size_t i = 0; i < listeners_.size(); i++)
            if (listeners_[i]->ID() == listener->ID()) {
                index = (int)i;
                break;
            }
        if (index == -1) return;
        ids_.Push(listener->ID());
        if (listeners_.size() > 1) listeners_[index] = listeners_.back();
        listeners_.pop_back();
    }
    const std::vector<Listener *> &Listeners() const { return listeners_; }

   private:
    int size_;
    std::vector<Listener *> listeners_;
    IDPool ids_;
    Factory factory_;
};

struct BlockReport {  // what the parity tests compare per block
    float psdNoiseFloor, noiseFloor, noiseDeviation, peakThreshold, listenThreshold;
    double noiseVariance;
};

// rx.Receiver (rx/receiver.go:64-464).  The goroutine + channel shell becomes an explicit queue: IQData enqueues
// (dropping when 100 frames wait, :328-333) and Process() runs the `case frame` body for everything queued, one
// GPU batch per cumulation window.
class Receiver {
   public:
    Receiver(const std::string &id, ReceiverMode mode, ManualClock *clock, sdr_engine *engine, int poolSize = defaultListenerPoolSize)
        : id_(id), mode_(mode), clock_(clock), engine_(engine),
          listeners_(mode == ReceiverMode::Decode ? 1 : poolSize, id, [this](const std::string &lid) { return newListener(lid); }) {}
    ~Receiver() { Stop(); }

    void AddReporter(Reporter *r) { reporter_ = r; }
    void Start(int sampleRate, int blockSize) {  // :130-146
        if (started_) return;
        sampleRate_ = sampleRate;
        blockSize_ = blockSize;
        frequencyMapping_.reset(new dsp::FrequencyMapping(sampleRate, blockSize, centerFrequency_));
        peaks_.reset(new PeaksTable(blockSize, clock_));
        peaks_->onDrop = [this](dsp::Peak *p) { releasePeak(p); };
        peaks_->deterministic = deterministicFindNext;
        peaks_->rng = rngSeed;
        if (sdr_stream_open(engine_, sampleRate, &stream_) != SDR_OK) throw std::runtime_error(sdr_last_error(engine_));
        void *p = nullptr;
        if (sdr_alloc_pinned(engine_, (size_t)cumulationSize * 2 * blockSize * sizeof(float), &p) != SDR_OK)
            throw std::runtime_error(sdr_last_error(engine_));
        ring_ = (float *)p;
        cumulationCount_ = 0;  // run() starts with cumulationCount := 0 (:347); sdr_stream_open reset the engine's count too
        started_ = true;
    }
    void Stop() {  // :148-164
        if (!started_) return;
        listeners_.Reset();
        std::fill(slots_.begin(), slots_.end(), nullptr);
        sdr_stream_close(engine_, stream_);
        sdr_free_pinned(engine_, ring_);
        ring_ = nullptr;
        started_ = false;
        queue_.clear();
    }
    void SetPeakThreshold(float t) { peakThreshold_ = t; }
    void SetEdgeWidth(int e) { edgeWidth_ = e; }
    void SetSilenceTimeout(int64_t t) { silenceTimeout_ = t; }
    void SetAttachmentTimeout(int64_t t) { attachmentTimeout_ = t; }
    void SetSignalDebounce(int d) { signalDebounce_ = d; }
    void SetCenterFrequency(int64_t f) {
        centerFrequency_ = f;
        if (frequencyMapping_) frequencyMapping_->SetCenterFrequency(f);
    }
    int64_t CenterFrequency() const { return centerFrequency_; }
    // decode mode / forced attach: rx/receiver.go:280-296 (ForcePut + Activate + Attach at a known bin)
    Listener *AttachAtBin(int bin) {
        Listener *l = listeners_.BindNext();
        if (!l) return nullptr;
        dsp::Peak peak = newPeakCenteredOnBin(bin);
        peak.SignalBin = bin;
        peak.SignalFrequency = frequencyMapping_->BinToFrequency(bin, dsp::BinCenter);
        peak.SignalValue = 80;
        dsp::Peak *pp = store(peak);
        peaks_->ForcePut(pp);
        peaks_->Activate(pp);
        l->Attach(pp);
        return l;
    }
    // :315-334: wrong rate / size are logged-and-dropped, a full queue drops the frame; returns false when dropped
    bool IQData(int sampleRate, const float *data, size_t len) {
        if (!started_) return false;
        if (sampleRate_ != sampleRate || blockSize_ != (int)(len / 2)) {
            rejected++;
            return false;
        }
        if ((int)queue_.size() >= iqBufferSize) {
            skipped++;
            return false;
        }
        queue_.emplace_back(data, data + len);
        return true;
    }
    // ---- batched interface: used by Process() below and by the Dispatcher (many receivers, one sdr_submit) ----
    // Moves the queued frames of this receiver -- at most up to the next flush: listeners change there -- into `dst`
    // (C-owned pinned memory, contiguous) and describes them as one sdr_work.  Returns the blocks staged (0: idle).
    int Stage(float *dst, sdr_work &w) {
        const int room = cumulationSize - cumulationCount_;
        const int nb = (int)std::min<size_t>(queue_.size(), (size_t)room);
        if (nb <= 0) return 0;
        for (int b = 0; b < nb; b++) {
            std::memcpy(dst + (size_t)b * 2 * blockSize_, queue_.front().data(), (size_t)2 * blockSize_ * sizeof(float));
            queue_.pop_front();
        }
        stagedBins_.clear();
        stagedFlags_.clear();
        if (deviceDebounce) {
            // one table entry per pool slot: a listener keeps its position for as long as it is bound, so the device can
            // carry its BoolDebouncer state from submit to submit
            stagedBins_.assign(slots_.size(), 0);
            stagedFlags_.assign(slots_.size(), 0);
            for (size_t sl = 0; sl < slots_.size(); sl++) {
                Listener *l = slots_[sl];
                if (!l) continue;
                l->tapIndex = l->Attached() ? (int)sl : -1;
                if (!l->Attached()) continue;
                stagedBins_[sl] = l->SignalBin();
                stagedFlags_[sl] = (uint8_t)(SDR_LISTENER_ACTIVE | (l->fresh ? SDR_LISTENER_RESET : 0));
                l->fresh = false;
            }
        } else {
            for (Listener *l : listeners_.Listeners()) {
                l->tapIndex = l->Attached() ? (int)stagedBins_.size() : -1;
                if (l->Attached()) stagedBins_.push_back(l->SignalBin());
            }
        }
        w = sdr_work{};
        w.stream = stream_;
        w.n_blocks = nb;
        w.iq = dst;
        w.mem = SDR_MEM_HOST;
        w.edge_width = edgeWidth_;
        w.peak_threshold = peakThreshold_;
        w.n_listeners = (int)stagedBins_.size();
        w.listener_bins = stagedBins_.data();
        w.signal_debounce = deviceDebounce ? signalDebounce_ : 1;
        w.listener_flags = deviceDebounce ? stagedFlags_.data() : nullptr;
        return nb;
    }
    // peaks are always scanned in strain mode: a listener may time out inside the batch and free a pool slot
    bool WantsPeaks() const { return mode_ == ReceiverMode::Strain; }
    // the rest of the frame iteration (:383-461) for work `wi` of a collected result
    void Consume(const sdr_result &r, int wi) {
        const int b0 = r.work_block_offset[wi], b1 = r.work_block_offset[wi + 1];
        for (int b = b0; b < b1; b++) consumeBlock(r, b);
        if (r.work_flush_offset[wi + 1] > r.work_flush_offset[wi]) consumeFlush(r, r.work_flush_offset[wi]);
    }
    bool Idle() const { return queue_.empty(); }
    bool Started() const { return started_; }
    int BlockSize() const { return blockSize_; }
    // The frame iteration of run() (:364-461) for every queued frame.  Returns the number of blocks processed.
    int Process() {
        int done = 0;
        sdr_work w;
        int nb;
        while ((nb = Stage(ring_, w)) > 0) {
            sdr_ticket t;
            if (sdr_submit(engine_, &w, 1, WantsPeaks() ? 0 : SDR_NO_PEAKS, &t) != SDR_OK) throw std::runtime_error(sdr_last_error(engine_));
            sdr_result r;
            if (sdr_collect(engine_, t, 1, &r) != SDR_OK) throw std::runtime_error(sdr_last_error(engine_));
            Consume(r, 0);
            sdr_release(engine_, t);
            done += nb;
        }
        return done;
    }

    bool recordReports = false;  // test hooks: keep one BlockReport per block / the peak list of every flush
    // run dsp.BoolDebouncer on the device: the host reads one packed, debounced bit per listener and block
    // and only ticks the decoder.  Set before Start().
    bool deviceDebounce = false;
    const std::vector<BlockReport> &Reports() const { return reports_; }
    const std::vector<std::unique_ptr<Listener>> &AllListeners() const { return allListeners_; }
    const std::vector<int64_t> &AttachBlocks() const { return attachBlocks_; }
    const std::vector<std::vector<dsp::Peak>> &FlushPeaks() const { return flushPeaks_; }
    ListenerPool &Pool() { return listeners_; }
    PeaksTable &Peaks() { return *peaks_; }
    int64_t BlocksProcessed() const { return blockIndex_; }
    int skipped = 0, rejected = 0;
    bool deterministicFindNext = true;
    uint64_t rngSeed = 1;
    bool blockClock = true;  // advance the manual clock to (b+1)*N/fs before block b is consumed

   private:
    Listener *newListener(const std::string &id) {  // :122-128
        allListeners_.emplace_back(new Listener(id, clock_, reporter_, sampleRate_, blockSize_));
        Listener *l = allListeners_.back().get();
        l->SetAttachmentTimeout(attachmentTimeout_);
        l->SetSilenceTimeout(silenceTimeout_);
        l->SetSignalDebounce(signalDebounce_);
        l->recordKeys = recordReports;  // the per-listener key log is a test hook like the block reports
        attachBlocks_.push_back(blockIndex_);
        if (slots_.empty()) slots_.assign((size_t)listeners_.Size(), nullptr);
        for (size_t sl = 0; sl < slots_.size(); sl++)
            if (!slots_[sl]) {
                slots_[sl] = l;
                l->slot = (int)sl;
                break;
            }
        return l;
    }
    dsp::Peak *store(const dsp::Peak &p) {
        peakStore_.emplace_back(new dsp::Peak(p));
        return peakStore_.back().get();
    }
    // the table dropped this peak; keep the storage only while a listener is still attached to it (a ForcePut can
    // replace an active peak, rx/peaks.go:46-60)
    void releasePeak(dsp::Peak *p) {
        for (Listener *l : listeners_.Listeners())
            if (l->Peak() == p) return;
        for (size_t k = 0; k < peakStore_.size(); k++)
            if (peakStore_[k].get() == p) {
                peakStore_[k].swap(peakStore_.back());
                peakStore_.pop_back();
                return;
            }
    }
    dsp::Peak newPeakCenteredOnBin(int centerBin) {  // :491-500
        dsp::Peak peak;
        peak.From = std::max(0, centerBin - peakPadding);
        peak.To = std::min(centerBin + peakPadding, blockSize_ - 1);
        peak.FromFrequency = frequencyMapping_->BinToFrequency(peak.From, dsp::BinFrom);
        peak.ToFrequency = frequencyMapping_->BinToFrequency(peak.To, dsp::BinTo);
        peak.SignalFrequency = peak.CenterFrequency();
        return peak;
    }
    void consumeBlock(const sdr_result &r, int b) {  // :379-407 minus everything the GPU did
        if (blockClock) clock_->Set((int64_t)(((__int128)(blockIndex_ + 1) * blockSize_ * kSecond) / sampleRate_));
        const int64_t nowS = clock_->Now() / kSecond;
        if (nowS > lastCleanupS_) {  // cleanupTicker, :359-363
            lastCleanupS_ = nowS;
            for (Listener *l : listeners_.Listeners()) l->CheckWriteTimeout();
            peaks_->Cleanup();
        }
        if (recordReports) {
            const float *th = r.thresholds + (size_t)b * 4;
            reports_.push_back(BlockReport{r.psd_noise_floor[b], th[0], th[1], th[2], th[3], r.noise_variance[b]});
        }
        std::vector<Listener *> detached;
        for (Listener *l : listeners_.Listeners()) {
            if (!l->Attached() || l->tapIndex < 0) continue;
            if (deviceDebounce) l->ListenDebounced(((r.key_bits[(size_t)b * r.key_words + (l->tapIndex >> 5)] >> (l->tapIndex & 31)) & 1u) != 0);
            else l->ListenState(r.keys[(size_t)b * r.tap_stride + l->tapIndex] != 0);
            if (mode_ == ReceiverMode::Strain && l->TimeoutExceeded()) {
                peaks_->Deactivate(l->Peak());
                l->Detach();
                detached.push_back(l);
            }
        }
        for (Listener *l : detached) {
            if (l->slot >= 0 && l->slot < (int)slots_.size()) slots_[l->slot] = nullptr;
            listeners_.Release(l);
        }
        cumulationCount_++;
        blockIndex_++;
    }
    void consumeFlush(const sdr_result &r, int f) {  // :409-460
        cumulationCount_ = 0;
        if (recordReports) flushPeaks_.emplace_back();
        if (mode_ != ReceiverMode::Strain || !listeners_.Available()) return;
        const int n = std::min(r.flush_n_peaks[f], r.max_peaks_per_flush);
        const sdr_peak *gp = r.flush_peaks + (size_t)f * r.max_peaks_per_flush;
        for (int i = 0; i < n; i++) {
            const dsp::Peak p = dsp::FromGpuPeak(gp[i], blockSize_, *frequencyMapping_);
            if (recordReports) flushPeaks_.back().push_back(p);
            dsp::Peak centered = newPeakCenteredOnBin(p.SignalBin);  // newPeakCenteredOnSignal :474-480
            centered.SignalFrequency = p.SignalFrequency;
            centered.SignalValue = p.SignalValue;
            centered.SignalBin = p.SignalBin;
            if (!peaks_->Put(store(centered))) peakStore_.pop_back();  // refused: nothing refers to it
        }
        dsp::Peak *selected = peaks_->FindNext();
        if (selected) {
            Listener *l = listeners_.BindNext();
            if (l) {
                attachBlocks_.back() = blockIndex_ - 1;
                peaks_->Activate(selected);
                l->Attach(selected);
            }
        }
    }

    std::string id_;
    ReceiverMode mode_;
    ManualClock *clock_;
    sdr_engine *engine_;
    Reporter *reporter_ = nullptr;
    float peakThreshold_ = defaultPeakThreshold;
    int edgeWidth_ = defaultEdgeWidth;
    int sampleRate_ = 0, blockSize_ = 0;
    int64_t centerFrequency_ = 0;
    int64_t silenceTimeout_ = defaultSilenceTimeout, attachmentTimeout_ = defaultAttachmentTimeout;
    int signalDebounce_ = 1;
    bool started_ = false;
    int stream_ = -1;
    float *ring_ = nullptr;
    std::vector<int> stagedBins_;
    std::vector<uint8_t> stagedFlags_;
    std::vector<Listener *> slots_;  // pool slot -> bound listener
    std::deque<std::vector<float>> queue_;
    std::unique_ptr<dsp::FrequencyMapping> frequencyMapping_;
    std::unique_ptr<PeaksTable> peaks_;
    ListenerPool listeners_;
    std::vector<std::unique_ptr<Listener>> allListeners_;
    std::vector<int64_t> attachBlocks_;
    std::vector<std::unique_ptr<dsp::Peak>> peakStore_;
    std::vector<BlockReport> reports_;
    std::vector<std::vector<dsp::Peak>> flushPeaks_;
    int cumulationCount_ = 0;
    int64_t blockIndex_ = 0, lastCleanupS_ = 0;
};

// Multi-receiver engine (rx/receiver.go:315-364 for R receivers at once): the reference runs one
// goroutine per rx.Receiver, each handling one frame at a time.  The GPU path wants the frames of ALL receivers that
// share an engine in ONE sdr_submit per tick -- one H2D copy (the works are staged back to back in one pinned arena),
// one K1 launch over every stream's segments, one K2, one D2H -- and hands each receiver its slice of the result.
// Receivers stay unchanged otherwise: IQData() enqueues (and drops when full), setters apply between ticks.
class Dispatcher {
   public:
    Dispatcher(sdr_engine *engine, int blockSize, int maxReceivers) : engine_(engine), blockSize_(blockSize) {
        void *p = nullptr;
        arenaFloats_ = (size_t)maxReceivers * cumulationSize * 2 * blockSize;
        if (sdr_alloc_pinned(engine_, arenaFloats_ * sizeof(float), &p) != SDR_OK) throw std::runtime_error(sdr_last_error(engine_));
        arena_ = (float *)p;
        maxReceivers_ = maxReceivers;
    }
    ~Dispatcher() { sdr_free_pinned(engine_, arena_); }
    void Add(Receiver *r) {
        if ((int)receivers_.size() >= maxReceivers_) throw std::invalid_argument("dispatcher is full");
        receivers_.push_back(r);
    }
    // Drains every receiver's queue.  One pass takes each receiver up to its next flush; passes repeat until all
    // queues are empty.  Returns the number of blocks processed; `submits` counts the sdr_submit calls made.
    int Tick() {
        int done = 0;
        for (;;) {
            works_.clear();
            owners_.clear();
            size_t off = 0;
            bool peaks = false;
            for (Receiver *r : receivers_) {
                if (!r->Started() || r->Idle() || r->BlockSize() != blockSize_) continue;
                sdr_work w;
                const int nb = r->Stage(arena_ + off, w);
                if (nb <= 0) continue;
                off += (size_t)nb * 2 * blockSize_;
                works_.push_back(w);
                owners_.push_back(r);
                peaks = peaks || r->WantsPeaks();
                done += nb;
            }
            if (works_.empty()) break;
            sdr_ticket t;
            if (sdr_submit(engine_, works_.data(), (int)works_.size(), peaks ? 0 : SDR_NO_PEAKS, &t) != SDR_OK)
                throw std::runtime_error(sdr_last_error(engine_));
            submits++;
            sdr_result r;
            if (sdr_collect(engine_, t, 1, &r) != SDR_OK) throw std::runtime_error(sdr_last_error(engine_));
            for (size_t i = 0; i < owners_.size(); i++) owners_[i]->Consume(r, (int)i);
            sdr_release(engine_, t);
        }
        return done;
    }
    int submits = 0;

   private:
    sdr_engine *engine_;
    int blockSize_, maxReceivers_ = 0;
    float *arena_ = nullptr;
    size_t arenaFloats_ = 0;
    std::vector<Receiver *> receivers_;
    std::vector<sdr_work> works_;
    std::vector<Receiver *> owners_;
};

}  // namespace rx
}  // namespace sdrhost
