// Package gp is the drop-in replacement of GoGP's package gp
// (bitbucket.org/dtolpin/gogp/gp, reference gp/gp.go) whose hot path runs on a
// B200 through the C-ABI of libgogp_b200.so (include/gogp_b200.h).
//
// The exported surface is the reference's: type Kernel, type GP with the fields
// NDim, Simil, Noise, ThetaSimil, ThetaNoise, X, Y, Parallel and the methods
// Absorb, LML, Produce, Observe, Gradient.  Observe+Gradient still satisfy
// Infergo's model.Model / ElementalModel, so the GP drops in under infer.FuncGrad,
// infer.Adam, HMC/NUTS and the tutorial models unchanged.
//
// NOT COMPILED IN THE BUILD ENVIRONMENT: no Go toolchain is available there; the
// executable counterpart that the parity tests drive is gogp_b200/gp.py (ctypes
// over the same entry points).  See INTEGRATION.md.
package gp

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gogp_b200 -lgogp_b200 -Wl,-rpath,${SRCDIR}/../../gogp_b200
#include <stdlib.h>
#include "gogp_b200.h"
*/
import "C"

import (
	"errors"
	"math"
	"runtime"
	"unsafe"

	"bitbucket.org/dtolpin/infergo/model"
)

// Kernel is the reference's kernel interface (gp/gp.go:14-17).
type Kernel interface {
	model.Model
	NTheta() int
}

// Op mirrors gogp_op: one instruction of the postfix kernel descriptor.
type Op struct {
	Kind     uint8
	Dim      uint8
	Param    [2]int16
	Scale    [2]float64
	Constant float64
}

// DeviceKernel is implemented by kernels that can run on the device: the stock
// kernels of package kernel and their sums / products / scalings.  A Simil that
// does not implement it cannot be used -- there is no CPU fallback.
type DeviceKernel interface {
	Kernel
	Descriptor() []Op
}

// GP is the reference's GP (gp/gp.go:20-38) with a device handle in place of
// the cached Cholesky factor.
type GP struct {
	NDim         int
	Simil, Noise Kernel

	ThetaSimil, ThetaNoise []float64
	X                      [][]float64
	Y                      []float64

	Parallel bool // kept for API parity; the GPU build is the parallel path
	Device   int  // CUDA device ordinal

	withObs bool
	n       int
	h       *C.gogp_handle
	key     [2]Kernel
}

func ops(k Kernel) ([]C.gogp_op, error) {
	dk, ok := k.(DeviceKernel)
	if !ok {
		return nil, errors.New("gp: kernel does not implement DeviceKernel; a host-language kernel cannot run on the device")
	}
	d := dk.Descriptor()
	out := make([]C.gogp_op, len(d))
	for i, o := range d {
		out[i].kind = C.uint8_t(o.Kind)
		out[i].dim = C.uint8_t(o.Dim)
		out[i].param[0], out[i].param[1] = C.int16_t(o.Param[0]), C.int16_t(o.Param[1])
		out[i].scale[0], out[i].scale[1] = C.double(o.Scale[0]), C.double(o.Scale[1])
		out[i].constant = C.double(o.Constant)
	}
	return out, nil
}

func (gp *GP) handle() (*C.gogp_handle, error) {
	if gp.h != nil && gp.key == [2]Kernel{gp.Simil, gp.Noise} {
		return gp.h, nil
	}
	gp.Close()
	so, err := ops(gp.Simil)
	if err != nil {
		return nil, err
	}
	var no []C.gogp_op
	ntn := 0
	if gp.Noise != nil { // nil -> the reference default ConstantNoise(1e-5), gp/gp.go:46-48
		if no, err = ops(gp.Noise); err != nil {
			return nil, err
		}
		ntn = gp.Noise.NTheta()
	}
	var h *C.gogp_handle
	var np *C.gogp_op
	if len(no) > 0 {
		np = &no[0]
	}
	st := C.gogp_create(C.int(gp.NDim), &so[0], C.int(len(so)), C.int(gp.Simil.NTheta()),
		np, C.int(len(no)), C.int(ntn), C.int(gp.Device), &h)
	if st != C.GOGP_OK {
		msg := C.GoString(C.gogp_last_error(h))
		C.gogp_destroy(h)
		return nil, errors.New("gp: " + msg)
	}
	gp.h, gp.key = h, [2]Kernel{gp.Simil, gp.Noise}
	runtime.SetFinalizer(gp, (*GP).Close)
	return h, nil
}

// SetEvents hands the event table of tutorial/events' Simil{Events} (rows of from, to, discount)
// to the device; used by kernel.Events leaves.
func (gp *GP) SetEvents(events [][]float64) error {
	h, err := gp.handle()
	if err != nil {
		return err
	}
	flat := flatten(events, 3)
	if st := C.gogp_set_events(h, dptr(flat), C.int(len(events))); st != C.GOGP_OK {
		return gp.fail(st)
	}
	return nil
}

// Close releases the device handle.
func (gp *GP) Close() {
	if gp.h != nil {
		C.gogp_destroy(gp.h)
		gp.h = nil
	}
}

func (gp *GP) ntn() int {
	if gp.Noise == nil {
		return 0
	}
	return gp.Noise.NTheta()
}

func (gp *GP) defaults() { // gp/gp.go:45-57
	if len(gp.ThetaSimil) == 0 {
		gp.ThetaSimil = make([]float64, gp.Simil.NTheta())
	}
	if len(gp.ThetaNoise) == 0 {
		gp.ThetaNoise = make([]float64, gp.ntn())
	}
}

// flatten copies [][]float64 into one contiguous row-major buffer: cgo may not
// be handed nested Go pointers, and C retains nothing after return.
func flatten(x [][]float64, ndim int) []float64 {
	out := make([]float64, 0, len(x)*ndim)
	for _, r := range x {
		out = append(out, r...)
	}
	return out
}

func dptr(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&s[0]))
}

func (gp *GP) fail(st C.gogp_status) error {
	return errors.New("gp: " + C.GoString(C.gogp_last_error(gp.h)))
}

// Absorb absorbs observations into the process (gp/gp.go:80-87).
func (gp *GP) Absorb(x [][]float64, y []float64) error {
	gp.defaults()
	gp.X, gp.Y = x, y
	h, err := gp.handle()
	if err != nil {
		return err
	}
	xf := flatten(x, gp.NDim)
	// the C-ABI takes one N for both buffers: a ragged x or len(x) != len(y) must not reach it
	if len(xf) != len(y)*gp.NDim {
		return errors.New("gp: len(x) != len(y) (or a row of x is not NDim long)")
	}
	gp.withObs, gp.n = false, len(y)
	if st := C.gogp_absorb(h, dptr(gp.ThetaSimil), dptr(gp.ThetaNoise), dptr(xf), dptr(y), C.int64_t(len(y))); st != C.GOGP_OK {
		return gp.fail(st)
	}
	return nil
}

// Extend appends observations to the absorbed ones at unchanged hyperparameters
// (the growing window of tutorial.Evaluate, tutorial/tutorial.go:91-179): the
// Cholesky factor is extended, O(N^2) per appended point, instead of recomputed.
// Equivalent to Absorb on the concatenated data.
func (gp *GP) Extend(x [][]float64, y []float64) error {
	h, err := gp.handle()
	if err != nil {
		return err
	}
	xf := flatten(x, gp.NDim)
	if len(xf) != len(y)*gp.NDim {
		return errors.New("gp: len(x) != len(y) (or a row of x is not NDim long)")
	}
	var lml C.double
	if st := C.gogp_extend(h, dptr(xf), dptr(y), C.int64_t(len(y)), &lml); st != C.GOGP_OK {
		return gp.fail(st)
	}
	gp.X = append(gp.X, x...)
	gp.Y = append(gp.Y, y...)
	gp.n += len(y)
	return nil
}

// LML is the log marginal likelihood of the absorbed observations (gp/gp.go:244-253).
func (gp *GP) LML() float64 {
	var out C.double
	if gp.h == nil || C.gogp_lml(gp.h, &out) != C.GOGP_OK {
		return 0
	}
	return float64(out)
}

// Produce computes predictions (gp/gp.go:258-360).
func (gp *GP) Produce(x [][]float64) (mu, sigma []float64, err error) {
	gp.defaults()
	h, err := gp.handle()
	if err != nil {
		return nil, nil, err
	}
	zf := flatten(x, gp.NDim)
	if len(zf) != len(x)*gp.NDim {
		return nil, nil, errors.New("gp: a row of x is not NDim long")
	}
	mu = make([]float64, len(x))
	sigma = make([]float64, len(x))
	if st := C.gogp_produce(h, dptr(zf), C.int64_t(len(x)), dptr(mu), dptr(sigma)); st != C.GOGP_OK {
		return nil, nil, gp.fail(st)
	}
	return mu, sigma, nil
}

// Observe computes the log marginal likelihood of the parameters given the
// observations (gp/gp.go:374-413).  The argument is the concatenation of
// log-transformed hyperparameters, inputs and outputs, or the hyperparameters
// only (then X, Y are the fields of gp).  As in the reference, the parameter
// prefix of x is exponentiated in place during the call and logged back, X and
// Y alias x in the first form, and failure panics.
func (gp *GP) Observe(x []float64) float64 {
	gp.defaults()
	h, err := gp.handle()
	if err != nil {
		panic(err)
	}
	nts, ntn := gp.Simil.NTheta(), gp.ntn()
	logTheta := append([]float64(nil), x[:nts+ntn]...)
	theta := x[:nts+ntn]
	for i := range theta {
		theta[i] = math.Exp(theta[i])
	}
	defer func() {
		for i := range theta {
			theta[i] = math.Log(theta[i])
		}
	}()
	copy(gp.ThetaSimil, model.Shift(&x, nts))
	copy(gp.ThetaNoise, model.Shift(&x, ntn))
	gp.withObs = len(x) > 0
	var xf, yf []float64
	if gp.withObs {
		n := len(x) / (gp.NDim + 1)
		xf = x[:n*gp.NDim]
		gp.X = make([][]float64, n)
		for i := range gp.X {
			gp.X[i] = model.Shift(&x, gp.NDim)
		}
		gp.Y = model.Shift(&x, n)
		yf = gp.Y
	} else {
		xf, yf = flatten(gp.X, gp.NDim), gp.Y
		if len(xf) != len(yf)*gp.NDim {
			panic("len(gp.X) != len(gp.Y)")
		}
	}
	if len(x) != 0 {
		panic("len(x)")
	}
	gp.n = len(yf)
	wo := C.int(0)
	if gp.withObs {
		wo = 1
	}
	var lml C.double
	if st := C.gogp_observe(h, dptr(logTheta), wo, dptr(xf), dptr(yf), C.int64_t(len(yf)), &lml); st != C.GOGP_OK {
		panic(gp.fail(st))
	}
	return float64(lml)
}

// Gradient computes the gradient of the log-likelihood with respect to the
// parameters and the inputs (gp/gp.go:418-499).
func (gp *GP) Gradient() []float64 {
	n := gp.Simil.NTheta() + gp.ntn()
	if gp.withObs {
		n += gp.n * (gp.NDim + 1)
	}
	grad := make([]float64, n)
	if gp.h == nil || n == 0 {
		return grad
	}
	if st := C.gogp_gradient(gp.h, dptr(grad), C.int64_t(n)); st != C.GOGP_OK {
		panic(gp.fail(st))
	}
	return grad
}

// Alpha is K^-1 y of the absorbed observations: the reference's exported gp.Alpha
// (gp/gp.go:36), which a user may store.
func (gp *GP) Alpha() []float64 {
	a := make([]float64, gp.n)
	if gp.h == nil || gp.n == 0 {
		return a
	}
	if st := C.gogp_get_alpha(gp.h, dptr(a), C.int64_t(gp.n)); st != C.GOGP_OK {
		panic(gp.fail(st))
	}
	return a
}

// L is the lower Cholesky factor of the covariance matrix, n x n row-major: the
// reference's exported gp.L (gp/gp.go:35).
func (gp *GP) L() []float64 {
	l := make([]float64, gp.n*gp.n)
	if gp.h == nil || gp.n == 0 {
		return l
	}
	if st := C.gogp_get_factor(gp.h, dptr(l), C.int64_t(gp.n)); st != C.GOGP_OK {
		panic(gp.fail(st))
	}
	return l
}

// Restore puts stored results back before Produce ("Produce works on stored
// results", gp/gp.go:255-257): the parameters and inputs of the stored
// factorisation with what Alpha() and L() returned.
func (gp *GP) Restore(thetaSimil, thetaNoise []float64, x [][]float64, alpha, l []float64) error {
	h, err := gp.handle()
	if err != nil {
		return err
	}
	xf := flatten(x, gp.NDim)
	n := len(alpha)
	if len(xf) != n*gp.NDim || len(l) != n*n || len(thetaSimil) != gp.Simil.NTheta() || len(thetaNoise) != gp.ntn() {
		return errors.New("gp: state shapes do not agree")
	}
	gp.ThetaSimil = append([]float64(nil), thetaSimil...)
	gp.ThetaNoise = append([]float64(nil), thetaNoise...)
	gp.X, gp.withObs, gp.n = x, false, n
	if st := C.gogp_set_state(h, dptr(gp.ThetaSimil), dptr(gp.ThetaNoise), dptr(xf), C.int64_t(n), dptr(alpha), dptr(l)); st != C.GOGP_OK {
		return gp.fail(st)
	}
	return nil
}

// OptResult reports what Optimize did (gogp_opt_result).
type OptResult struct {
	Iters, Evals, Grads int
	LML0, LML           float64
	Converged           bool
}

// Optimize runs the tutorial's MLE loop (tutorial/tutorial.go:124-175) inside
// the library over gp.X, gp.Y: alg "adam" is the infer.Adam loop, "lbfgs" the
// optimize.Minimize call; iters, threshold, rate are ITERS, THRESHOLD, RATE.  x
// holds the log hyperparameters and is updated in place.  This form maximises
// the marginal likelihood alone; priors (gp.Model) go through gogp_optimize's
// callback argument with an exported cgo function -- see INTEGRATION.md.
func (gp *GP) Optimize(x []float64, alg string, iters int, threshold, rate float64) (OptResult, error) {
	gp.defaults()
	h, err := gp.handle()
	if err != nil {
		return OptResult{}, err
	}
	if len(x) != gp.Simil.NTheta()+gp.ntn() {
		return OptResult{}, errors.New("len(x)")
	}
	xf := flatten(gp.X, gp.NDim)
	if len(xf) != len(gp.Y)*gp.NDim {
		return OptResult{}, errors.New("gp: len(gp.X) != len(gp.Y)")
	}
	if st := C.gogp_set_data(h, dptr(xf), dptr(gp.Y), C.int64_t(len(gp.Y))); st != C.GOGP_OK {
		return OptResult{}, gp.fail(st)
	}
	var s C.gogp_opt_settings
	if alg == "lbfgs" {
		s.method = 1
	}
	s.max_iters = C.int(iters)
	s.threshold = C.double(threshold)
	s.rate = C.double(rate)
	var r C.gogp_opt_result
	if st := C.gogp_optimize(h, &s, dptr(x), nil, nil, &r); st != C.GOGP_OK {
		return OptResult{}, gp.fail(st)
	}
	gp.withObs, gp.n = false, len(gp.Y)
	nts := gp.Simil.NTheta()
	for i := range x {
		if i < nts {
			gp.ThetaSimil[i] = math.Exp(x[i])
		} else {
			gp.ThetaNoise[i-nts] = math.Exp(x[i])
		}
	}
	return OptResult{int(r.iters), int(r.evals), int(r.grads), float64(r.lml0), float64(r.lml), r.converged != 0}, nil
}
