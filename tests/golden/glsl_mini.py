"""A minimal CPU evaluator for the GLSL subset the reference's hex-mosaic fragment shader uses
(HyGrid/HexPixelArt/hexagon_mosaic_shader.py:26-82): float / int / vec2 variables, arithmetic, ``int()`` / ``float()`` /
``vec2()`` constructors, ``& % == <``, ``if / else`` blocks, ``texture2D``.  TEST INFRASTRUCTURE for
tests/golden/make_mosaic_golden.py: the build container has no OpenGL, so the shader TEXT of the reference is translated
statement by statement into Python and executed per fragment with GLSL's value semantics:

* ``float`` is IEEE binary32 (numpy.float32), every operation rounded to it;
* ``int`` is a 32-bit integer: ``/`` truncates toward zero, ``%`` takes the sign of the dividend (GLSL 3.30, section 5.9), ``int(float)`` truncates;
* an int operand meeting a float operand is converted to float (implicit conversion, section 4.1.10).

The program that runs is the reference's; only these three rules are ours."""
import re

import numpy as np

f32 = np.float32


class I:
    def __init__(self, v):
        self.v = int(v)

    @staticmethod
    def lift(o):
        return o if isinstance(o, (I, F)) else (I(o) if isinstance(o, int) else F(o))

    def _f(self):
        return F(f32(self.v))

    def _bin(self, o, fi, ff):
        o = I.lift(o)
        return fi(self.v, o.v) if isinstance(o, I) else ff(self._f(), o)

    def __add__(self, o): return self._bin(o, lambda a, b: I(a + b), lambda a, b: a + b)
    def __radd__(self, o): return I.lift(o) + self
    def __sub__(self, o): return self._bin(o, lambda a, b: I(a - b), lambda a, b: a - b)
    def __rsub__(self, o): return I.lift(o) - self
    def __mul__(self, o): return self._bin(o, lambda a, b: I(a * b), lambda a, b: a * b)
    def __rmul__(self, o): return I.lift(o) * self

    def __truediv__(self, o):
        def idiv(a, b):
            q = abs(a) // abs(b)
            return I(q if (a >= 0) == (b >= 0) else -q)
        return self._bin(o, idiv, lambda a, b: a / b)

    def __rtruediv__(self, o): return I.lift(o) / self

    def __mod__(self, o):
        o = I.lift(o)
        a, b = self.v, o.v
        r = abs(a) % abs(b)
        return I(r if a >= 0 else -r)

    def __and__(self, o): return I(self.v & I.lift(o).v)
    def __eq__(self, o): return self.v == I.lift(o).v
    def __lt__(self, o): return self._bin(o, lambda a, b: a < b, lambda a, b: a < b)


class F:
    def __init__(self, v):
        self.v = f32(v.v if isinstance(v, (I, F)) else v)

    @staticmethod
    def of(o):
        return o if isinstance(o, F) else F(o)

    def __add__(self, o): return F(self.v + F.of(o).v)
    def __radd__(self, o): return F.of(o) + self
    def __sub__(self, o): return F(self.v - F.of(o).v)
    def __rsub__(self, o): return F.of(o) - self
    def __mul__(self, o): return F(self.v * F.of(o).v)
    def __rmul__(self, o): return F.of(o) * self
    def __truediv__(self, o): return F(self.v / F.of(o).v)
    def __rtruediv__(self, o): return F.of(o) / self
    def __lt__(self, o): return bool(self.v < F.of(o).v)


class Vec2:
    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = F.of(x), F.of(y)


def to_int(v):
    return I(int(np.trunc(v.v))) if isinstance(v, F) else I(v.v if isinstance(v, I) else v)


def to_float(v):
    return F(v)


def translate(body: str) -> str:
    """GLSL statements of ``main`` -> Python source (one statement per line, blocks by indentation)."""
    body = re.sub(r"//[^\n]*", "", body)
    out, depth = [], 1
    # split on ; { } keeping the block structure
    for tok in re.findall(r"[^;{}]+[;{]|}", body):
        tok = tok.strip()
        if tok == "}":
            depth -= 1
            continue
        opener = tok.endswith("{")
        stmt = tok[:-1].strip()
        if not stmt and opener:
            depth += 1
            continue
        m = re.match(r"^(if)\s*\((.*)\)$", stmt, flags=re.S)
        if m:
            out.append("    " * depth + f"if {expr(m.group(2))}:")
        elif stmt == "else":
            out.append("    " * depth + "else:")
        else:
            m = re.match(r"^(float|int|vec2|vec4)\s+(.*)$", stmt, flags=re.S)
            if m:
                kind, rest = m.groups()
                if "=" in rest:
                    name, rhs = rest.split("=", 1)
                    conv = {"float": "to_float", "int": "to_int"}.get(kind)
                    out.append("    " * depth + f"{name.strip()} = " + (f"{conv}({expr(rhs)})" if conv else expr(rhs)))
                else:
                    for name in rest.split(","):
                        init = {"float": "F(0.0)", "int": "I(0)"}.get(kind, "Vec2()")
                        out.append("    " * depth + f"{name.strip()} = {init}")
            else:
                name, rhs = stmt.split("=", 1)
                out.append("    " * depth + f"{name.strip()} = {expr(rhs)}")
        if opener:
            depth += 1
            out.append("    " * depth + "pass")
    return "\n".join(out)


def expr(e: str) -> str:
    e = e.strip()
    e = re.sub(r"\bint\s*\(", "to_int(", e)
    e = re.sub(r"\bfloat\s*\(", "to_float(", e)
    e = re.sub(r"\bvec2\s*\(", "Vec2(", e)
    # literals: 0.5 -> F(0.5), 1 -> I(1)   (identifiers containing digits are left alone)
    e = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)(?![\w.])", r"F(\1)", e)
    e = re.sub(r"(?<![\w.(])(\d+)(?![\w.])", r"I(\1)", e)
    return e


def compile_fragment_shader(src: str):
    """``run(uv_x, uv_y, size_x, size_y, ratio, even_odd_offset) -> (sx, sy)``: the texture coordinate (in texels) the
    shader samples -- everything up to the ``texture2D`` call of the reference's ``main``."""
    main = src[src.index("void main()"):]
    body = main[main.index("{") + 1:main.rindex("}")]
    body = body[:body.index("vec4 color")]                       # the fetch itself needs a texture unit
    py = translate(body)
    code = ("def run(uvx, uvy, sizex_in, sizey_in, ratio, eoo):\n"
            "    uv = Vec2(uvx, uvy); size = Vec2(sizex_in, sizey_in); hexmosaicSizeRatio = F(ratio); even_odd_offset = I(eoo)\n"
            "    sizex = F(0.0); sizey = F(0.0)\n" + py + "\n    return sx, sy\n")
    ns = dict(I=I, F=F, Vec2=Vec2, to_int=to_int, to_float=to_float)
    exec(compile(code, "hexagon_mosaic_shader.fs", "exec"), ns)
    return ns["run"], code
