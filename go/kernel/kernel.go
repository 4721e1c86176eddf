// Package kernel is the device-descriptor counterpart of GoGP's package kernel
// (reference kernel/kernel.go, kernel/noise.go): the same singletons Normal,
// Periodic, Matern32, Matern52, ConstantNoise, UniformNoise, which additionally
// implement gp.DeviceKernel, plus combinators (Param, Const, Sum, Prod, On) that
// replace hand-written Observe methods such as tutorial/*/kernel/kernel.go.
//
// Observe stays available (it is the reference's CPU arithmetic, useful for unit
// tests of user code) but gp.GP never calls it: the device evaluates the
// descriptor.  NOT COMPILED IN THE BUILD ENVIRONMENT (no Go toolchain).
package kernel

import (
	"math"

	"bitbucket.org/dtolpin/gogp-b200/go/gp"
)

// gogp_op_kind
const (
	opConst = iota
	opParam
	opAdd
	opMul
	opNormal
	opPeriodic
	opMatern32
	opMatern52
	opMatern52Textbook
	opEvents
)

// Expr is a kernel given by its postfix descriptor.
type Expr struct {
	ops    []gp.Op
	ntheta int
}

func (e Expr) Descriptor() []gp.Op { return e.ops }
func (e Expr) NTheta() int         { return e.ntheta }

// Observe and Gradient make Expr a model.Model; they are not on the GPU path.
func (e Expr) Observe(x []float64) float64 { return evalHost(e.ops, x, e.ntheta) }
func (e Expr) Gradient() []float64         { return nil }

// WithNTheta declares more parameters than the expression uses
// (tutorial/anynoise/kernel/kernel.go:31-35).
func (e Expr) WithNTheta(n int) Expr { return Expr{e.ops, n} }

func maxInt(a, b int) int {
	if a > b {
		return a
	}
	return b
}

func leaf(kind uint8, l, p int, ls, ps float64, dim int, nparam int) Expr {
	return Expr{[]gp.Op{{Kind: kind, Dim: uint8(dim), Param: [2]int16{int16(l), int16(p)}, Scale: [2]float64{ls, ps}}},
		maxInt(l, p*(nparam-1)) + 1}
}

// Stock kernels with the reference's parameter layout ([l] or [l, p] at 0, 1).
var (
	Normal   = leaf(opNormal, 0, 0, 1, 1, 0, 1)   // kernel/kernel.go:13-26
	Periodic = leaf(opPeriodic, 0, 1, 1, 1, 0, 2) // kernel/kernel.go:34-47
	Matern32 = leaf(opMatern32, 0, 0, 1, 1, 0, 1) // kernel/kernel.go:60-73
	Matern52 = leaf(opMatern52, 0, 0, 1, 1, 0, 1) // kernel/kernel.go:79-92 (5/3 == 1 as shipped)
	// UniformNoise: variance theta[0]^2 (kernel/noise.go:39-53)
	UniformNoise = Prod(Param(0, 1), Param(0, 1))
)

// On re-targets a stock kernel: length scale (and period) slots with constant
// multipliers, acting on input coordinate dim.
func On(k Expr, dim, l int, lscale float64, p int, pscale float64) Expr {
	o := k.ops[0]
	o.Dim = uint8(dim)
	o.Param = [2]int16{int16(l), int16(p)}
	o.Scale = [2]float64{lscale, pscale}
	n := l + 1
	if o.Kind == opPeriodic {
		n = maxInt(l, p) + 1
	}
	return Expr{[]gp.Op{o}, n}
}

func Param(i int, scale float64) Expr {
	return Expr{[]gp.Op{{Kind: opParam, Param: [2]int16{int16(i), 0}, Scale: [2]float64{scale, 1}}}, i + 1}
}

// Events is the discount factor of tutorial/events/kernel/kernel.go:33-44 on input coordinate
// dim; the event table itself is handed to the device with gp.GP.SetEvents (gogp_set_events).
func Events(dim int) Expr { return Expr{[]gp.Op{{Kind: opEvents, Dim: uint8(dim)}}, 0} }

func Const(c float64) Expr { return Expr{[]gp.Op{{Kind: opConst, Constant: c}}, 0} }

// ConstantNoise is kernel.ConstantNoise (kernel/noise.go:21-34): variance std^2.
func ConstantNoise(std float64) Expr { return Const(std * std) }

func bin(kind uint8, a, b Expr) Expr {
	ops := append(append(append([]gp.Op{}, a.ops...), b.ops...), gp.Op{Kind: kind})
	return Expr{ops, maxInt(a.ntheta, b.ntheta)}
}

func Sum(a, b Expr) Expr  { return bin(opAdd, a, b) }
func Prod(a, b Expr) Expr { return bin(opMul, a, b) }

// evalHost evaluates the descriptor with the reference's arithmetic for
// x = [theta | xa | xb]; host-side convenience only.
func evalHost(ops []gp.Op, x []float64, ntheta int) float64 {
	ndim := (len(x) - ntheta) / 2
	var st []float64
	for _, o := range ops {
		var xa, xb float64
		if o.Kind >= opNormal && ndim > 0 {
			xa, xb = x[ntheta+int(o.Dim)], x[ntheta+ndim+int(o.Dim)]
		}
		l := o.Scale[0]
		if o.Kind == opParam || o.Kind >= opNormal {
			l *= x[o.Param[0]]
		}
		switch o.Kind {
		case opConst:
			st = append(st, o.Constant)
		case opParam:
			st = append(st, l)
		case opAdd, opMul:
			a, b := st[len(st)-2], st[len(st)-1]
			st = st[:len(st)-2]
			if o.Kind == opAdd {
				st = append(st, a+b)
			} else {
				st = append(st, a*b)
			}
		case opNormal:
			d := (xa - xb) / l
			st = append(st, math.Exp(-d*d/2))
		case opPeriodic:
			d := math.Sin(math.Pi*math.Abs(xa-xb)/(o.Scale[1]*x[o.Param[1]])) / l
			st = append(st, math.Exp(-2*d*d))
		case opMatern32:
			d := math.Abs(xa-xb) / l
			st = append(st, (1+1.7320508075688772*d)*math.Exp(-1.7320508075688772*d))
		case opMatern52, opMatern52Textbook:
			c := 1.0
			if o.Kind == opMatern52Textbook {
				c = 5.0 / 3.0
			}
			d := math.Abs(xa-xb) / l
			st = append(st, (1+2.2360679774997900*d+c*d*d)*math.Exp(-2.2360679774997900*d))
		case opEvents:
			// the event table lives with the GP (gp.GP.SetEvents), not with the expression: the
			// host-side convenience evaluates the undiscounted kernel
			st = append(st, 1)
		}
	}
	return st[0]
}
