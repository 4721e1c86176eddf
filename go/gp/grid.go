package gp

/*
#include <stdlib.h>
#include "gogp_b200.h"
*/
import "C"

import (
	"errors"
	"runtime"
)

// Grid is a GP whose covariance matrix is spread over the GPUs of one box: K is
// dealt 2D block-cyclically to a Pr x Pc grid of devices and Observe / Gradient
// (reference gp/gp.go:374-413, 418-499) run as a block Cholesky plus one fused
// pass for K^-1, with the panel exchange (NCCL / NVLink peer memory) inside the
// library (gogp_b200/csrc/grid.hpp, grid.cu).  One Go process, one handle, one
// cgo call per Observe: the library runs a host thread per device for the
// duration of a call.  Hyper-parameters-only mode: X and Y are set once with
// SetData (the reference assigns gp.X, gp.Y: tutorial/tutorial.go:114-115).
//
// Observe + Gradient satisfy Infergo's model.Model, so infer.FuncGrad,
// optimize.Minimize and infer.Adam drive a Grid as they drive a GP.
//
// NOT COMPILED IN THE BUILD ENVIRONMENT (no Go toolchain); the executable mirror
// is gogp_b200/grid.py (GridGP), which the GPU tests and bench.py drive.
type Grid struct {
	NDim         int
	Simil, Noise Kernel

	ThetaSimil, ThetaNoise []float64

	g *C.gogp_grid
	n int
}

// NewGrid forms the grid on the given CUDA devices.  pr * pc must equal
// len(devices); pr = pc = 0 selects the default shape (pr >= pc, powers of two).
func NewGrid(ndim int, simil, noise Kernel, devices []int, pr, pc int) (*Grid, error) {
	so, err := ops(simil)
	if err != nil {
		return nil, err
	}
	var no []C.gogp_op
	ntn := 0
	if noise != nil {
		if no, err = ops(noise); err != nil {
			return nil, err
		}
		ntn = noise.NTheta()
	}
	var np *C.gogp_op
	if len(no) > 0 {
		np = &no[0]
	}
	if len(devices) == 0 {
		return nil, errors.New("gp: a grid needs at least one device")
	}
	devs := make([]C.int, len(devices))
	for i, d := range devices {
		devs[i] = C.int(d)
	}
	var g *C.gogp_grid
	st := C.gogp_create_grid(C.int(ndim), &so[0], C.int(len(so)), C.int(simil.NTheta()), np, C.int(len(no)), C.int(ntn),
		&devs[0], C.int(len(devs)), C.int(pr), C.int(pc), 0, &g)
	if st != C.GOGP_OK {
		msg := C.GoString(C.gogp_grid_last_error(g))
		C.gogp_grid_destroy(g)
		return nil, errors.New("gp: " + msg)
	}
	gr := &Grid{NDim: ndim, Simil: simil, Noise: noise, g: g}
	runtime.SetFinalizer(gr, (*Grid).Close)
	return gr, nil
}

// Close releases the devices' memory and the communicator.
func (g *Grid) Close() {
	if g.g != nil {
		C.gogp_grid_destroy(g.g)
		g.g = nil
	}
}

func (g *Grid) ntn() int {
	if g.Noise == nil {
		return 0
	}
	return g.Noise.NTheta()
}

func (g *Grid) fail() error {
	return errors.New("gp: " + C.GoString(C.gogp_grid_last_error(g.g)))
}

// SetData replicates the observations on every device of the grid and allocates
// each device's share of K (about N^2 * 8 / len(devices) bytes).
func (g *Grid) SetData(x [][]float64, y []float64) error {
	xf := flatten(x, g.NDim)
	if len(xf) != len(y)*g.NDim {
		return errors.New("gp: len(x) != len(y) (or a row of x is not NDim long)")
	}
	if st := C.gogp_grid_set_data(g.g, dptr(xf), dptr(y), C.int64_t(len(y))); st != C.GOGP_OK {
		return g.fail()
	}
	g.n = len(y)
	return nil
}

// Observe is GP.Observe in the hyper-parameters-only form: x holds the log
// hyper-parameters; panics where the reference panics (gp/gp.go:398-405).
func (g *Grid) Observe(x []float64) float64 {
	if len(x) != g.Simil.NTheta()+g.ntn() {
		panic("len(x)")
	}
	var lml C.double
	if st := C.gogp_grid_observe(g.g, dptr(x), &lml); st != C.GOGP_OK {
		panic(g.fail())
	}
	return float64(lml)
}

// Gradient is GP.Gradient for the last Observe (gp/gp.go:418-499).
func (g *Grid) Gradient() []float64 {
	grad := make([]float64, g.Simil.NTheta()+g.ntn())
	if len(grad) == 0 {
		return grad
	}
	if st := C.gogp_grid_gradient(g.g, dptr(grad), C.int64_t(len(grad))); st != C.GOGP_OK {
		panic(g.fail())
	}
	return grad
}

// Absorb is GP.Absorb (gp/gp.go:80-87) with the natural-scale parameters of
// ThetaSimil / ThetaNoise over the data of SetData; LML follows.
func (g *Grid) Absorb() error {
	if len(g.ThetaSimil) == 0 {
		g.ThetaSimil = make([]float64, g.Simil.NTheta())
	}
	if len(g.ThetaNoise) == 0 {
		g.ThetaNoise = make([]float64, g.ntn())
	}
	if st := C.gogp_grid_absorb(g.g, dptr(g.ThetaSimil), dptr(g.ThetaNoise)); st != C.GOGP_OK {
		return g.fail()
	}
	return nil
}

// LML is GP.LML (gp/gp.go:244-253).
func (g *Grid) LML() float64 {
	var out C.double
	if g.g == nil || C.gogp_grid_lml(g.g, &out) != C.GOGP_OK {
		return 0
	}
	return float64(out)
}

// Alpha is K^-1 y (available after Gradient).
func (g *Grid) Alpha() []float64 {
	a := make([]float64, g.n)
	if g.n == 0 {
		return a
	}
	if st := C.gogp_grid_get_alpha(g.g, dptr(a), C.int64_t(g.n)); st != C.GOGP_OK {
		panic(g.fail())
	}
	return a
}
