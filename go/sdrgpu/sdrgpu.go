//go:build sdrgpu

// Package sdrgpu binds libsdrgpu.so (include/sdrgpu.h) for ftl/sdrainer.
//
// UNCOMPILED: there is no Go toolchain in the build image of this repository, so this file has never
// been through `go build`/`go vet`.  It is the binding a maintainer adds next to rx/receiver.go; the C ABI
// itself is exercised by the Python/ctypes and C++ tests of this repository.
package sdrgpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../sdrainer_b200 -lsdrgpu -Wl,-rpath,${SRCDIR}/../../sdrainer_b200
#include <stdlib.h>
#include "sdrgpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"unsafe"
)

// Engine owns one GPU engine.  The C ABI serialises concurrent callers internally, so several receiver goroutines may
// share one Engine; a stream and a ticket still have one owner at a time (like rx.Receiver.run owns its state).
type Engine struct {
	h         *C.sdr_engine
	blockSize int
}

type Config struct {
	Device, BlockSize, MaxStreams, MaxListeners, MaxBlocksPerBatch, MaxPeaksPerFlush, Slots int
}

func New(cfg Config) (*Engine, error) {
	c := C.sdr_engine_config{
		device: C.int(cfg.Device), block_size: C.int(cfg.BlockSize), max_streams: C.int(cfg.MaxStreams),
		max_listeners: C.int(cfg.MaxListeners), max_blocks_per_batch: C.int(cfg.MaxBlocksPerBatch),
		max_peaks_per_flush: C.int(cfg.MaxPeaksPerFlush), n_slots: C.int(cfg.Slots),
	}
	var h *C.sdr_engine
	if rc := C.sdr_engine_create(&c, &h); rc != C.SDR_OK {
		return nil, fmt.Errorf("sdrgpu: %s", C.GoString(C.sdr_last_error(nil)))
	}
	return &Engine{h: h, blockSize: cfg.BlockSize}, nil
}

func (e *Engine) Close() { C.sdr_engine_destroy(e.h); e.h = nil }

func (e *Engine) err(rc C.int) error {
	if rc == C.SDR_OK {
		return nil
	}
	if rc == C.SDR_EBUSY {
		return ErrBusy
	}
	return errors.New(C.GoString(C.sdr_last_error(e.h)))
}

// ErrBusy means "no free slot": the caller drops the frames, as rx.Receiver.IQData does when r.in is full
// (rx/receiver.go:328-333).
var ErrBusy = errors.New("sdrgpu: all in-flight slots busy")

// PinnedFloats returns C-owned pinned memory viewed as a Go slice.  cgo forbids C from retaining Go pointers
// after a call returns, so the IQ ring lives on the C side and Go writes into it.
func (e *Engine) PinnedFloats(n int) ([]float32, error) {
	var p unsafe.Pointer
	if err := e.err(C.sdr_alloc_pinned(e.h, C.size_t(n*4), &p)); err != nil {
		return nil, err
	}
	return unsafe.Slice((*float32)(p), n), nil
}

func (e *Engine) OpenStream(sampleRate int) (int, error) {
	var s C.int
	err := e.err(C.sdr_stream_open(e.h, C.int(sampleRate), &s))
	return int(s), err
}

// Work mirrors C.sdr_work: the frames currently queued in r.in for one receiver.
type Work struct {
	Stream        int
	IQ            []float32 // slice of PinnedFloats memory, len = nBlocks*2*blockSize
	EdgeWidth     int
	PeakThreshold float32
	ListenerBins  []int32
	// SignalDebounce is SpectralDemodulator.SetSignalDebounce (cw/spectral.go:33-35); the BoolDebouncer of every listener
	// position runs on the device and Result.Key returns the debounced state.  0 or 1 = pass-through.
	SignalDebounce int
	// ListenerFlags (optional, one per bin): FlagActive for a position with an attached listener, | FlagReset when a new
	// listener was bound to the position since the previous submit.  nil = all active, none reset.  Keep a listener at
	// the same position (its pool slot) for as long as it is bound.
	ListenerFlags []uint8
}

const (
	FlagActive = uint8(C.SDR_LISTENER_ACTIVE)
	FlagReset  = uint8(C.SDR_LISTENER_RESET)
	// submit flags
	NoTaps    = int(C.SDR_NO_TAPS)
	NoRawKeys = int(C.SDR_NO_RAW_KEYS)
	NoPeaks   = int(C.SDR_NO_PEAKS)
)

type Ticket int64

func (e *Engine) Submit(works []Work, flags int) (Ticket, error) {
	cw := make([]C.sdr_work, len(works))
	pins := make([]unsafe.Pointer, 0, len(works))
	defer func() {
		for _, p := range pins {
			C.free(p)
		}
	}()
	for i, w := range works {
		cw[i].stream = C.int(w.Stream)
		cw[i].n_blocks = C.int(len(w.IQ) / (2 * e.blockSize))
		cw[i].iq = (*C.float)(unsafe.Pointer(&w.IQ[0]))
		cw[i].mem = C.SDR_MEM_HOST
		cw[i].format = C.SDR_FMT_F32
		cw[i].edge_width = C.int(w.EdgeWidth)
		cw[i].peak_threshold = C.float(w.PeakThreshold)
		cw[i].n_listeners = C.int(len(w.ListenerBins))
		if len(w.ListenerBins) > 0 {
			// listener bins are copied into C memory: the C side reads them only during this call
			p := C.malloc(C.size_t(len(w.ListenerBins) * 4))
			pins = append(pins, p)
			copy(unsafe.Slice((*int32)(p), len(w.ListenerBins)), w.ListenerBins)
			cw[i].listener_bins = (*C.int)(p)
		}
		cw[i].signal_debounce = C.int(w.SignalDebounce)
		if len(w.ListenerFlags) == len(w.ListenerBins) && len(w.ListenerFlags) > 0 {
			p := C.malloc(C.size_t(len(w.ListenerFlags)))
			pins = append(pins, p)
			copy(unsafe.Slice((*uint8)(p), len(w.ListenerFlags)), w.ListenerFlags)
			cw[i].listener_flags = (*C.uint8_t)(p)
		}
	}
	var t C.sdr_ticket
	err := e.err(C.sdr_submit(e.h, &cw[0], C.int(len(cw)), C.int(flags), &t))
	return Ticket(t), err
}

// Result is a view into engine-owned pinned host memory, valid until Release.
type Result struct{ r C.sdr_result }

func (e *Engine) Collect(t Ticket) (*Result, error) {
	res := &Result{}
	err := e.err(C.sdr_collect(e.h, C.sdr_ticket(t), 1, &res.r))
	return res, err
}

func (e *Engine) Release(t Ticket) error { return e.err(C.sdr_release(e.h, C.sdr_ticket(t))) }

func (r *Result) Blocks() int { return int(r.r.n_blocks) }

// ListenThreshold is noiseFloor+noiseDeviation of block b (rx/receiver.go:394).
func (r *Result) ListenThreshold(b int) float32 {
	return unsafe.Slice((*float32)(unsafe.Pointer(r.r.thresholds)), 4*int(r.r.n_blocks))[4*b+3]
}

// PeakThreshold is r.peakThreshold+noiseFloor of block b (rx/receiver.go:385).
func (r *Result) PeakThreshold(b int) float32 {
	return unsafe.Slice((*float32)(unsafe.Pointer(r.r.thresholds)), 4*int(r.r.n_blocks))[4*b+2]
}

// Tap is spectrum[l.SignalBin()] of listener l in block b (rx/receiver.go:393).
func (r *Result) Tap(b, l int) float32 {
	s := int(r.r.tap_stride)
	return unsafe.Slice((*float32)(unsafe.Pointer(r.r.taps)), s*int(r.r.n_blocks))[b*s+l]
}

// Key is the debounced key state of listener position l in block b (cw/spectral.go:48-50): what cw.Decoder.Tick takes.
func (r *Result) Key(b, l int) bool {
	kw := int(r.r.key_words)
	words := unsafe.Slice((*uint32)(unsafe.Pointer(r.r.key_bits)), kw*int(r.r.n_blocks))
	return (words[b*kw+l/32]>>(uint(l)%32))&1 != 0
}

// WorkBlocks returns the block range [from, to) of work w inside this result (several receivers per submit).
func (r *Result) WorkBlocks(w int) (int, int) {
	off := unsafe.Slice((*C.int)(unsafe.Pointer(r.r.work_block_offset)), int(r.r.n_works)+1)
	return int(off[w]), int(off[w+1])
}

// WorkFlushes returns the flush range [from, to) of work w.
func (r *Result) WorkFlushes(w int) (int, int) {
	off := unsafe.Slice((*C.int)(unsafe.Pointer(r.r.work_flush_offset)), int(r.r.n_works)+1)
	return int(off[w]), int(off[w+1])
}

// Peaks returns the dsp.FindPeaks list of flush f (bin order).
func (r *Result) Peaks(f int) []C.sdr_peak {
	mp := int(r.r.max_peaks_per_flush)
	n := int(unsafe.Slice((*C.int)(unsafe.Pointer(r.r.flush_n_peaks)), int(r.r.n_flushes))[f])
	if n > mp {
		n = mp
	}
	all := unsafe.Slice((*C.sdr_peak)(unsafe.Pointer(r.r.flush_peaks)), mp*int(r.r.n_flushes))
	return all[f*mp : f*mp+n]
}

// Source is what the dispatcher needs from a receiver: rx.Receiver implements it with the frames queued in r.in
// (rx/receiver.go:315-334) and the loop body of run() after the FFT (rx/receiver.go:383-461).
type Source interface {
	// Stage copies the queued frames -- at most up to the next flush, listeners change there -- into dst (pinned
	// memory) and describes them; it returns false when nothing is queued.
	Stage(dst []float32) (Work, bool)
	// Consume handles the receiver's slice of a result: work index w of res.
	Consume(res *Result, w int)
}

// Dispatcher drains the queues of many receivers into ONE Submit per tick: one H2D copy (the works are staged back to
// back in one pinned arena), one spectral launch over every stream, one result.  It is the Go counterpart of
// rx::Dispatcher in sdrainer_b200/host/sdrhost.hpp (tested there against 64 oracle receivers).
type Dispatcher struct {
	eng     *Engine
	arena   []float32
	sources []Source
}

func NewDispatcher(eng *Engine, maxReceivers int) (*Dispatcher, error) {
	arena, err := eng.PinnedFloats(maxReceivers * 100 * 2 * eng.blockSize)
	if err != nil {
		return nil, err
	}
	return &Dispatcher{eng: eng, arena: arena}, nil
}

func (d *Dispatcher) Add(s Source) { d.sources = append(d.sources, s) }

// Tick processes everything queued; it returns the number of Submit calls it made.
func (d *Dispatcher) Tick() (int, error) {
	submits := 0
	for {
		works := make([]Work, 0, len(d.sources))
		owners := make([]Source, 0, len(d.sources))
		off := 0
		for _, s := range d.sources {
			w, ok := s.Stage(d.arena[off:])
			if !ok {
				continue
			}
			off += len(w.IQ)
			works = append(works, w)
			owners = append(owners, s)
		}
		if len(works) == 0 {
			return submits, nil
		}
		t, err := d.eng.Submit(works, NoTaps|NoRawKeys)
		if err != nil {
			return submits, err
		}
		submits++
		res, err := d.eng.Collect(t)
		if err != nil {
			return submits, err
		}
		for i, s := range owners {
			s.Consume(res, i)
		}
		if err := d.eng.Release(t); err != nil {
			return submits, err
		}
	}
}
