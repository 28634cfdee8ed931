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

// Engine owns one GPU engine; like rx.Receiver.run it must be used from one goroutine.
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
}

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
