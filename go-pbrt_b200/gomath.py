"""Go's math.Sin/Cos/Tan (golang:1.11 src/math/{sin,tan}.go, Cephes-derived) for the HOST-side transform constructors
(RotateX/Y/Z, Perspective — pkg/pbrt/transform.go:381-424,492-502).  Go's trig is not bit-identical to libm
(pkg/pbrt/transform_test.go:77-81 expects cos(Pi/2) = 6.123233995736757e-17), and the matrices the library receives
must be the ones the Go host would build.  Python float arithmetic is IEEE double without FMA contraction."""
import math

Pi = math.pi
PI4A = 7.85398125648498535156e-1
PI4B = 3.77489470793079817668e-8
PI4C = 2.69515142907905952645e-15
M4PI = 1.273239544735162542821171882678754627704620361328125  # Go 1.11 literal (0x1.45f306dc9c882p+0)

_sin = (1.58962301576546568060e-10, -2.50507477628578072866e-8, 2.75573136213857245213e-6,
        -1.98412698295895385996e-4, 8.33333333332211858878e-3, -1.66666666666666307295e-1)
_cos = (-1.13585365213876817300e-11, 2.08757008419747316778e-9, -2.75573141792967388112e-7,
        2.48015872888517045348e-5, -1.38888888888730564116e-3, 4.16666666666665929218e-2)
_tanP = (-1.30936939181383777646e4, 1.15351664838587416140e6, -1.79565251976484877988e7)
_tanQ = (1.0, 1.36812963470692954678e4, -1.32089234440210967447e6, 2.50083801823357915839e7, -5.38695755929454629881e7)


def Radians(deg):  # pkg/math/math.go:130-132
    return Pi / 180.0 * deg


def _psin(z, zz):
    return z + z * zz * ((((((_sin[0] * zz) + _sin[1]) * zz + _sin[2]) * zz + _sin[3]) * zz + _sin[4]) * zz + _sin[5])


def _pcos(zz):
    return 1.0 - 0.5 * zz + zz * zz * ((((((_cos[0] * zz) + _cos[1]) * zz + _cos[2]) * zz + _cos[3]) * zz + _cos[4]) * zz + _cos[5])


def Cos(x):
    if math.isnan(x) or math.isinf(x):
        return math.nan
    sign = False
    x = abs(x)
    j = int(x * M4PI)
    y = float(j)
    if j & 1:
        j += 1
        y += 1
    j &= 7
    if j > 3:
        j -= 4
        sign = not sign
    if j > 1:
        sign = not sign
    z = ((x - y * PI4A) - y * PI4B) - y * PI4C
    zz = z * z
    y = _psin(z, zz) if j in (1, 2) else _pcos(zz)
    return -y if sign else y


def Sin(x):
    if x == 0 or math.isnan(x):
        return x
    if math.isinf(x):
        return math.nan
    sign = False
    if x < 0:
        x, sign = -x, True
    j = int(x * M4PI)
    y = float(j)
    if j & 1:
        j += 1
        y += 1
    j &= 7
    if j > 3:
        sign = not sign
        j -= 4
    z = ((x - y * PI4A) - y * PI4B) - y * PI4C
    zz = z * z
    y = _pcos(zz) if j in (1, 2) else _psin(z, zz)
    return -y if sign else y


def Tan(x):
    if x == 0 or math.isnan(x):
        return x
    if math.isinf(x):
        return math.nan
    sign = False
    if x < 0:
        x, sign = -x, True
    j = int(x * M4PI)
    y = float(j)
    if j & 1:
        j += 1
        y += 1
    z = ((x - y * PI4A) - y * PI4B) - y * PI4C
    zz = z * z
    if zz > 1e-14:
        y = z + z * (zz * (((_tanP[0] * zz) + _tanP[1]) * zz + _tanP[2]) / ((((zz + _tanQ[1]) * zz + _tanQ[2]) * zz + _tanQ[3]) * zz + _tanQ[4]))
    else:
        y = z
    if j & 2 == 2:
        y = -1 / y
    return -y if sign else y
