// host_capi.cpp -- C entry points over sdrhost.hpp so that the pytest suite can drive the C++ host mirror
// (the tests mirror cw/decode_test.go, dsp/dsp_test.go, dsp/fft_test.go, rx/peaks_test.go, rx/listener_test.go).
#include <cstdio>
#include <string>
#include <vector>

#include "sdrhost.hpp"
#include "realtime.hpp"

using namespace sdrhost;

namespace {
struct StringWriter : cw::Writer {
    std::string text;
    void Write(const std::string &s) override { text += s; }
};
struct DecoderBox {
    StringWriter out;
    cw::Decoder dec;
    DecoderBox(int fs, int bs) : dec(&out, fs, bs) {}
};
struct PeaksBox {
    rx::ManualClock clock;
    rx::PeaksTable table;
    std::vector<std::unique_ptr<dsp::Peak>> peaks;
    explicit PeaksBox(int size) : table(size, &clock) { clock.Set(1000 * rx::kSecond); }
    int id_of(const dsp::Peak *p) const {
        for (size_t i = 0; i < peaks.size(); i++)
            if (peaks[i].get() == p) return (int)i;
        return -1;
    }
};
struct PoolBox {
    rx::ManualClock clock;
    std::vector<std::unique_ptr<rx::Listener>> made;
    rx::ListenerPool pool;
    PoolBox(int size, const char *prefix)
        : pool(size, prefix, [this](const std::string &id) {
              made.emplace_back(new rx::Listener(id, &clock, nullptr, 48000, 512));
              return made.back().get();
          }) {}
};
struct EventLog : rx::Reporter {
    std::vector<std::string> events;
    void ListenerActivated(const std::string &l, int64_t f) override { events.push_back("+" + l + "@" + std::to_string(f)); }
    void ListenerDeactivated(const std::string &l, int64_t f) override { events.push_back("-" + l + "@" + std::to_string(f)); }
};
struct ReceiverBox {
    rx::ManualClock clock;
    EventLog log;
    rx::Receiver rx;
    std::string scratch;
    ReceiverBox(sdr_engine *e, int strain, int pool) : rx("rx", strain ? rx::ReceiverMode::Strain : rx::ReceiverMode::Decode, &clock, e, pool) {
        rx.AddReporter(&log);
        rx.recordReports = true;  // the parity tests read one report per block and every flush's peak list
    }
};
struct DispatcherBox {
    rx::Dispatcher d;
    std::string scratch;
    DispatcherBox(sdr_engine *e, int bs, int max_rx) : d(e, bs, max_rx) {}
};
}  // namespace

extern "C" {

// ---- cw.Decoder ----
void *sdrh_decoder_new(int fs, int bs) { return new DecoderBox(fs, bs); }
void sdrh_decoder_free(void *p) { delete (DecoderBox *)p; }
void sdrh_decoder_reset(void *p) { ((DecoderBox *)p)->dec.Reset(); }
void sdrh_decoder_tick(void *p, int s) { ((DecoderBox *)p)->dec.Tick(s != 0); }
void sdrh_decoder_ticks(void *p, const unsigned char *s, int n) {
    for (int i = 0; i < n; i++) ((DecoderBox *)p)->dec.Tick(s[i] != 0);
}
void sdrh_decoder_stop(void *p) { ((DecoderBox *)p)->dec.stop(); }
const char *sdrh_decoder_text(void *p) { return ((DecoderBox *)p)->out.text.c_str(); }
void sdrh_decoder_clear_text(void *p) { ((DecoderBox *)p)->out.text.clear(); }

// ---- dsp ----
void *sdrh_debouncer_new(int thr) { return new dsp::BoolDebouncer(thr); }
void sdrh_debouncer_free(void *p) { delete (dsp::BoolDebouncer *)p; }
int sdrh_debouncer_debounce(void *p, int raw) { return ((dsp::BoolDebouncer *)p)->Debounce(raw != 0) ? 1 : 0; }
long long sdrh_bin_to_frequency(int fs, int bs, long long center, int bin, double location) {
    return dsp::FrequencyMapping(fs, bs, center).BinToFrequency(bin, location);
}
int sdrh_frequency_to_bin(int fs, int bs, long long center, long long frequency) {
    return dsp::FrequencyMapping(fs, bs, center).FrequencyToBin(frequency);
}
long long sdrh_peak_signal_frequency(int fs, int bs, long long center, int bin, float y1, float y2, float y3) {
    sdr_peak g{bin, bin, bin, 0.f, y1, y2, y3};
    return dsp::FromGpuPeak(g, bs, dsp::FrequencyMapping(fs, bs, center)).SignalFrequency;
}

// ---- rx.PeaksTable ----
void *sdrh_peaks_new(int size) { return new PeaksBox(size); }
void sdrh_peaks_free(void *p) { delete (PeaksBox *)p; }
int sdrh_peaks_make(void *p, int from, int to) {
    PeaksBox *b = (PeaksBox *)p;
    b->peaks.emplace_back(new dsp::Peak());
    b->peaks.back()->From = from;
    b->peaks.back()->To = to;
    return (int)b->peaks.size() - 1;
}
void sdrh_peaks_put(void *p, int id, int force) {
    PeaksBox *b = (PeaksBox *)p;
    if (force) b->table.ForcePut(b->peaks[id].get());
    else b->table.Put(b->peaks[id].get());
}
void sdrh_peaks_activate(void *p, int id) { ((PeaksBox *)p)->table.Activate(((PeaksBox *)p)->peaks[id].get()); }
void sdrh_peaks_deactivate(void *p, int id) { ((PeaksBox *)p)->table.Deactivate(((PeaksBox *)p)->peaks[id].get()); }
void sdrh_peaks_cleanup(void *p) { ((PeaksBox *)p)->table.Cleanup(); }
void sdrh_peaks_clock_add(void *p, double seconds) { ((PeaksBox *)p)->clock.Add((int64_t)(seconds * 1e9)); }
int sdrh_peaks_find_next(void *p) {
    PeaksBox *b = (PeaksBox *)p;
    return b->id_of(b->table.FindNext());
}
int sdrh_peaks_bin(void *p, int bin) {  // peak id occupying the bin or -1
    PeaksBox *b = (PeaksBox *)p;
    return b->id_of(b->table.Get(bin));
}
int sdrh_peaks_bin_state(void *p, int bin) {
    const rx::PeaksTable::Entry *e = ((PeaksBox *)p)->table.Bin(bin);
    return e ? (int)e->state : 0;
}

// ---- rx.IDPool / rx.ListenerPool ----
void *sdrh_pool_new(int size, const char *prefix) { return new PoolBox(size, prefix); }
void sdrh_pool_free(void *p) { delete (PoolBox *)p; }
int sdrh_pool_bind_next(void *p) {  // index into the made[] list or -1
    PoolBox *b = (PoolBox *)p;
    rx::Listener *l = b->pool.BindNext();
    if (!l) return -1;
    for (size_t i = 0; i < b->made.size(); i++)
        if (b->made[i].get() == l) return (int)i;
    return -1;
}
const char *sdrh_pool_made_id(void *p, int i) { return ((PoolBox *)p)->made[i]->ID().c_str(); }
void sdrh_pool_release(void *p, int i) { ((PoolBox *)p)->pool.Release(((PoolBox *)p)->made[i].get()); }
int sdrh_pool_len(void *p) { return (int)((PoolBox *)p)->pool.Listeners().size(); }
const char *sdrh_pool_active_id(void *p, int slot) { return ((PoolBox *)p)->pool.Listeners()[slot]->ID().c_str(); }

// ---- rx.Receiver over a GPU engine ----
void *sdrh_receiver_new(void *engine, int strain, int pool_size) { return new ReceiverBox((sdr_engine *)engine, strain, pool_size); }
void sdrh_receiver_free(void *p) { delete (ReceiverBox *)p; }
int sdrh_receiver_start(void *p, int fs, int bs) {
    try {
        ((ReceiverBox *)p)->rx.Start(fs, bs);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
void sdrh_receiver_stop(void *p) { ((ReceiverBox *)p)->rx.Stop(); }
void sdrh_receiver_set(void *p, float peak_threshold, int edge_width, double silence_s, double attach_s, int debounce, long long center) {
    rx::Receiver &r = ((ReceiverBox *)p)->rx;
    r.SetPeakThreshold(peak_threshold);
    r.SetEdgeWidth(edge_width);
    r.SetSilenceTimeout((int64_t)(silence_s * 1e9));
    r.SetAttachmentTimeout((int64_t)(attach_s * 1e9));
    r.SetSignalDebounce(debounce);
    r.SetCenterFrequency(center);
}
// rx/peaks.go:183-207: FindNext's random probe, seeded (the reference uses unseeded math/rand); call before start
void sdrh_receiver_set_find_next(void *p, int deterministic, unsigned long long seed) {
    rx::Receiver &r = ((ReceiverBox *)p)->rx;
    r.deterministicFindNext = deterministic != 0;
    r.rngSeed = seed;
}
void sdrh_receiver_set_device_debounce(void *p, int on) { ((ReceiverBox *)p)->rx.deviceDebounce = on != 0; }
int sdrh_receiver_iq_data(void *p, int fs, const float *data, long long len) { return ((ReceiverBox *)p)->rx.IQData(fs, data, (size_t)len) ? 1 : 0; }
int sdrh_receiver_process(void *p) {
    try {
        return ((ReceiverBox *)p)->rx.Process();
    } catch (const std::exception &e) {
        ((ReceiverBox *)p)->scratch = e.what();
        return -1;
    }
}
const char *sdrh_receiver_error(void *p) { return ((ReceiverBox *)p)->scratch.c_str(); }
int sdrh_receiver_attach_at_bin(void *p, int bin) { return ((ReceiverBox *)p)->rx.AttachAtBin(bin) ? 0 : -1; }
int sdrh_receiver_skipped(void *p) { return ((ReceiverBox *)p)->rx.skipped; }
int sdrh_receiver_rejected(void *p) { return ((ReceiverBox *)p)->rx.rejected; }
int sdrh_receiver_listener_count(void *p) { return (int)((ReceiverBox *)p)->rx.AllListeners().size(); }
int sdrh_receiver_listener_bin(void *p, int i) {
    rx::Listener *l = ((ReceiverBox *)p)->rx.AllListeners()[i].get();
    return l->Attached() ? l->SignalBin() : -1;
}
const char *sdrh_receiver_listener_text(void *p, int i) { return ((ReceiverBox *)p)->rx.AllListeners()[i]->Text().c_str(); }
const unsigned char *sdrh_receiver_listener_keys(void *p, int i, long long *n) {
    const std::vector<uint8_t> &k = ((ReceiverBox *)p)->rx.AllListeners()[i]->Keys();
    *n = (long long)k.size();
    return k.data();
}
long long sdrh_receiver_attach_block(void *p, int i) { return ((ReceiverBox *)p)->rx.AttachBlocks()[i]; }
int sdrh_receiver_n_reports(void *p) { return (int)((ReceiverBox *)p)->rx.Reports().size(); }
void sdrh_receiver_report(void *p, int b, float *out5, double *var) {
    const rx::BlockReport &r = ((ReceiverBox *)p)->rx.Reports()[b];
    out5[0] = r.psdNoiseFloor;
    out5[1] = r.noiseFloor;
    out5[2] = r.noiseDeviation;
    out5[3] = r.peakThreshold;
    out5[4] = r.listenThreshold;
    *var = r.noiseVariance;
}
int sdrh_receiver_n_events(void *p) { return (int)((ReceiverBox *)p)->log.events.size(); }
const char *sdrh_receiver_event(void *p, int i) { return ((ReceiverBox *)p)->log.events[i].c_str(); }
int sdrh_receiver_n_flushes(void *p) { return (int)((ReceiverBox *)p)->rx.FlushPeaks().size(); }
int sdrh_receiver_flush_peaks(void *p, int f, int *bins, long long *freqs, int cap) {
    const auto &v = ((ReceiverBox *)p)->rx.FlushPeaks()[f];
    int n = 0;
    for (const auto &pk : v) {
        if (n < cap) {
            bins[n] = pk.SignalBin;
            freqs[n] = pk.SignalFrequency;
        }
        n++;
    }
    return n;
}

// ---- rx.Dispatcher: many receivers, one sdr_submit per tick ----
void *sdrh_dispatcher_new(void *engine, int block_size, int max_receivers) {
    try {
        return new DispatcherBox((sdr_engine *)engine, block_size, max_receivers);
    } catch (const std::exception &) {
        return nullptr;
    }
}
void sdrh_dispatcher_free(void *p) { delete (DispatcherBox *)p; }
int sdrh_dispatcher_add(void *p, void *receiver) {
    try {
        ((DispatcherBox *)p)->d.Add(&((ReceiverBox *)receiver)->rx);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
int sdrh_dispatcher_tick(void *p) {
    try {
        return ((DispatcherBox *)p)->d.Tick();
    } catch (const std::exception &e) {
        ((DispatcherBox *)p)->scratch = e.what();
        return -1;
    }
}
int sdrh_dispatcher_submits(void *p) { return ((DispatcherBox *)p)->d.submits; }
const char *sdrh_dispatcher_error(void *p) { return ((DispatcherBox *)p)->scratch.c_str(); }

// ---- real-time channel-count harness (realtime.hpp) ----
void *sdrh_rt_new(void *engine, int fs, int n, int listeners, int s_cap, int blocks_per_batch, int threads, int debounce, const float *src,
                  int n_templates, int src_blocks, const int *bins) {
    try {
        return new rt::Harness((sdr_engine *)engine, fs, n, listeners, s_cap, blocks_per_batch, threads, debounce, src, n_templates,
                               src_blocks, bins);
    } catch (const std::exception &) {
        return nullptr;
    }
}
void sdrh_rt_free(void *p) { delete (rt::Harness *)p; }
// out[0..8] = batch_s, copy_s, submit_s, collect_wait_s, decode_s, gpu_ms, ticks, chars, key_downs
int sdrh_rt_run(void *p, int n_streams, int n_batches, int ring_copy, double *out) {
    try {
        const rt::Stats st = ((rt::Harness *)p)->Run(n_streams, n_batches, ring_copy != 0);
        out[0] = st.batch_s;
        out[1] = st.copy_s;
        out[2] = st.submit_s;
        out[3] = st.collect_wait_s;
        out[4] = st.decode_s;
        out[5] = st.gpu_ms;
        out[6] = (double)st.ticks;
        out[7] = (double)st.chars;
        out[8] = (double)st.key_downs;
        return 0;
    } catch (const std::exception &ex) {
        fprintf(stderr, "sdrh_rt_run: %s\n", ex.what());  // the engine's message (sdr_last_error) for the caller's log
        return -1;
    }
}

// ---- cw.AudioDemodulator over the GPU Goertzel bank ----
struct AudioBox {
    StringWriter out;
    cw::AudioDemodulator dem;
    AudioBox(double pitch, int fs) : dem(&out, pitch, fs) {}
};
void *sdrh_audio_new(double pitch, int fs) {
    try {
        return new AudioBox(pitch, fs);
    } catch (const std::exception &) {
        return nullptr;
    }
}
void sdrh_audio_free(void *p) { delete (AudioBox *)p; }
int sdrh_audio_blocksize(void *p) { return ((AudioBox *)p)->dem.Blocksize(); }
void sdrh_audio_set_scale(void *p, double s) { ((AudioBox *)p)->dem.SetScale(s); }
int sdrh_audio_write(void *p, const float *buf, int n) {
    try {
        return ((AudioBox *)p)->dem.Write(buf, n);
    } catch (const std::exception &) {
        return -1;
    }
}
void sdrh_audio_close(void *p) { ((AudioBox *)p)->dem.Close(); }
const char *sdrh_audio_text(void *p) { return ((AudioBox *)p)->out.text.c_str(); }

}  // extern "C"
