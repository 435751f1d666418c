"""
TEST INFRASTRUCTURE ONLY (oracle). Never imported by the product path; only `tests/`, the golden generators
under `tests/golden/` and `bench.py`'s CPU legs may use it.

A mechanical evaluator for the GLSL 3.30 subset the reference's in-scope shaders are written in. It is NOT a
restatement of any shader: it is fed the shader TEXT the reference hands to the OpenGL driver
(`ShaderProgram.compile`, shaderflow/shader.py:313-324 — header assembly `_build_shader` :190-239 over
`resources/shaders/include/shaderflow.glsl`, `include/camera.glsl`, `vertex/default.glsl`,
`fragment/final.glsl`, `examples/**.frag`; captured by `oracle/ref_scene.py` from the reference's own Python)
and executes it SIMT-style over numpy arrays, one array lane per fragment:

    preprocess()   #define (object / function-like, `##`, line continuation), #ifdef/#ifndef/#else/#endif,
                   #undef; comments stripped
    Parser         declarations (const / uniform / in / out / flat), structs, overloaded functions, arrays,
                   if / for / while / do / switch / break / continue / return / discard, the full expression
                   grammar (assignment operators, ?:, swizzles, field and index access, constructors)
    Machine        typed values (float / int / uint / bool, vec*, ivec*, bvec*, mat2-4, structs, arrays,
                   sampler2D), masked execution for divergent control flow, user-function calls with
                   in / out / inout parameters, GLSL's implicit int→float conversions (incl. the NVIDIA-style
                   leniencies the reference relies on)

What remains a RESTATEMENT of a specification rather than of the reference (the list the oracle's users must
know; everything else comes from the shader text itself):
  * arithmetic is IEEE float32 with one rounding per operation, no FMA contraction, no compile-time folding in
    higher precision (float loop counters therefore run in strict float32: visualizer.frag:26 → 9 directions);
  * builtin functions follow GLSL 3.30 §8 literally: mix = x·(1−a)+y·a, clamp = min(max(x,lo),hi),
    smoothstep = t²(3−2t), mod = x−y·floor(x/y), length = sqrt(dot), normalize = v/length, dot and matrix
    products summed left to right; min / max / clamp drop NaNs like the hardware's FMNMX (fminf / fmaxf);
    transcendental functions are numpy's float32 ufuncs (within a few ulp; GLSL allows far more);
  * int(x) truncates toward zero; integer division truncates toward zero;
  * texture() / texelFetch() / textureSize(): OpenGL 3.3 fixed-function sampling, restated in
    `oracle/glsl_np.Texture` (pixel centres, bilinear weights, wrap modes, unorm8 decode);
  * rasterisation of the fullscreen quad: varyings are the exact affine interpolation of the VERTEX SHADER's
    outputs (the vertex shader text is executed by this evaluator on the four quad vertices) at the fragment
    centres, computed in float64 and rounded once to float32; colour stores to 8-bit targets clamp and
    round-half-even (`glsl_np.to_unorm8`).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

F32, I32, U32 = np.float32, np.int32, np.uint32

# ---------------------------------------------------------------------------------------------- #
# Preprocessor

_TOKEN = re.compile(r"""
    (?P<float>(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?[fF]?|\d+[eE][+-]?\d+[fF]?)
  | (?P<int>0[xX][0-9a-fA-F]+[uU]?|\d+[uU]?)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op>\#\#|\+\+|--|<<=|>>=|<<|>>|<=|>=|==|!=|&&|\|\||\^\^|\+=|-=|\*=|/=|%=|&=|\|=|\^=|[-+*/%<>=!&|^~?:;,.(){}\[\]\#])
  | (?P<ws>\s+)
""", re.VERBOSE)


def tokenize(text: str) -> list[tuple[str, str]]:
    out, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise SyntaxError(f"GLSL: cannot tokenize at {text[pos:pos+30]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind != "ws":
            out.append((kind, m.group()))
    return out


def strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", lambda m: "\n"*m.group().count("\n"), text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def text_digest(text: str) -> str:
    """sha1 of the token stream (comments and whitespace do not count: the assembled header carries module
    uuids in its section comments, shader.py:203-229)"""
    import hashlib
    return hashlib.sha1(" ".join(v for _, v in tokenize(strip_comments(text).replace("\\\n", " "))).encode()).hexdigest()


@dataclass
class Macro:
    params: list[str] | None
    body: list[tuple[str, str]]


def preprocess(text: str, defines: dict[str, str] | None = None) -> list[tuple[str, str]]:
    """→ token list of the translation unit after directive processing and macro expansion"""
    text = strip_comments(text).replace("\\\n", " ")
    macros: dict[str, Macro] = {k: Macro(None, tokenize(v)) for k, v in (defines or {}).items()}
    out: list[tuple[str, str]] = []
    pending: list[tuple[str, str]] = []    # text tokens not yet macro-expanded (macros apply from their #define on)
    stack: list[list[bool]] = []           # [active, taken, parent_active]
    active = True

    def flush():
        out.extend(_expand(pending, macros, frozenset()))
        pending.clear()
    for line in text.split("\n"):
        s = line.strip()
        if s.startswith("#"):
            m = re.match(r"#\s*(\w+)\s*(.*)", s)
            if not m:
                continue
            d, rest = m.group(1), m.group(2).strip()
            if d in ("ifdef", "ifndef"):
                cond = (rest.split()[0] in macros) == (d == "ifdef")
                stack.append([active and cond, cond, active])
                active = stack[-1][0]
            elif d == "if":
                cond = _pp_condition(rest, macros)
                stack.append([active and cond, cond, active])
                active = stack[-1][0]
            elif d == "elif":
                top = stack[-1]
                cond = (not top[1]) and _pp_condition(rest, macros)
                top[0] = top[2] and cond
                top[1] = top[1] or cond
                active = top[0]
            elif d == "else":
                top = stack[-1]
                top[0] = top[2] and not top[1]
                top[1] = True
                active = top[0]
            elif d == "endif":
                active = stack.pop()[2]
            elif not active:
                continue
            elif d == "define":
                flush()
                mm = re.match(r"(\w+)(\(([^)]*)\))?\s*(.*)", rest, re.S)
                name, has_params, params, body = mm.group(1), mm.group(2), mm.group(3), mm.group(4)
                # a function-like macro needs '(' directly after the name
                if has_params and rest[len(name):len(name) + 1] == "(":
                    plist = [p.strip() for p in params.split(",") if p.strip()]
                    macros[name] = Macro(plist, tokenize(body))
                else:
                    macros[name] = Macro(None, tokenize(rest[len(name):]))
            elif d == "undef":
                flush()
                macros.pop(rest.split()[0], None)
            elif d in ("version", "extension", "pragma", "line"):
                continue
            else:
                raise SyntaxError(f"GLSL: unsupported directive #{d}")
            continue
        if active and s:
            pending.extend(tokenize(line))
    flush()
    return out


def _pp_condition(expr: str, macros) -> bool:
    expr = re.sub(r"defined\s*\(\s*(\w+)\s*\)|defined\s+(\w+)",
                  lambda m: "1" if (m.group(1) or m.group(2)) in macros else "0", expr)
    toks = _expand(tokenize(expr), macros, frozenset())
    src = " ".join(v for _, v in toks).replace("&&", " and ").replace("||", " or ").replace("!", " not ")
    src = re.sub(r"\bnot\s*=", "!=", src)
    src = re.sub(r"[A-Za-z_]\w*", lambda m: m.group() if m.group() in ("and", "or", "not") else "0", src)
    return bool(eval(src, {"__builtins__": {}}))


def _expand(tokens, macros, hidden):
    out, i = [], 0
    while i < len(tokens):
        kind, val = tokens[i]
        mac = macros.get(val) if kind == "id" and val not in hidden else None
        if mac is None:
            out.append(tokens[i]); i += 1
            continue
        if mac.params is None:
            out.extend(_expand(mac.body, macros, hidden | {val}))
            i += 1
            continue
        if i + 1 >= len(tokens) or tokens[i + 1][1] != "(":
            out.append(tokens[i]); i += 1
            continue
        args, depth, cur, j = [], 0, [], i + 2
        while True:
            k, v = tokens[j]
            if v == "(":
                depth += 1
            elif v == ")":
                if depth == 0:
                    break
                depth -= 1
            if v == "," and depth == 0:
                args.append(cur); cur = []
            else:
                cur.append(tokens[j])
            j += 1
        if cur or args:
            args.append(cur)
        bind = dict(zip(mac.params, args))
        body, b, n = [], 0, len(mac.body)
        while b < n:
            k, v = mac.body[b]
            pasting = (b + 1 < n and mac.body[b + 1][1] == "##") or (b > 0 and mac.body[b - 1][1] == "##")
            if v == "##":
                left = body.pop()
                rk, rv = mac.body[b + 1]
                right = bind[rv] if (rk == "id" and rv in bind) else [(rk, rv)]
                glued = tokenize(left[1] + (right[0][1] if right else ""))
                body.extend(glued); body.extend(right[1:])
                b += 2
                continue
            if k == "id" and v in bind:
                body.extend(bind[v] if pasting else _expand(bind[v], macros, hidden))
            else:
                body.append((k, v))
            b += 1
        out.extend(_expand(body, macros, hidden | {val}))
        i = j + 1
    return out

# ---------------------------------------------------------------------------------------------- #
# Types

SCALARS = {"float": F32, "int": I32, "uint": U32, "bool": np.bool_}
VEC_BASE = {"vec": "float", "ivec": "int", "uvec": "uint", "bvec": "bool"}
BASIC_TYPES = set(SCALARS) | {f"{p}{n}" for p in VEC_BASE for n in (2, 3, 4)} | {"mat2", "mat3", "mat4", "void",
                                                                                 "sampler2D"}
QUALIFIERS = {"const", "in", "out", "inout", "uniform", "flat", "smooth", "noperspective", "highp", "mediump",
              "lowp", "precise", "centroid", "varying", "attribute"}


def vec_info(t: str):
    """→ (base scalar type, dimension) for scalars and vectors, None otherwise"""
    if t in SCALARS:
        return t, 1
    for p, b in VEC_BASE.items():
        if t.startswith(p) and t[len(p):] in ("2", "3", "4") and (p != "vec" or not t.startswith("ivec")):
            return b, int(t[len(p):])
    return None


def make_type(base: str, n: int) -> str:
    if n == 1:
        return base
    return {"float": "vec", "int": "ivec", "uint": "uvec", "bool": "bvec"}[base] + str(n)


class V:
    """A typed SIMT value. Scalars: a.shape = (L,); vectors (L, n); matrices (L, cols, rows), L ∈ {1, lanes}.
    Structs: a = {field: V}; arrays: a = [V, ...]; sampler2D: a = Texture"""
    __slots__ = ("t", "a")

    def __init__(self, t, a):
        self.t, self.a = t, a

    def __repr__(self):
        return f"V({self.t}, {getattr(self.a, 'shape', self.a)})"

# ---------------------------------------------------------------------------------------------- #
# Parser → tuples ("kind", ...)

class Parser:
    def __init__(self, tokens):
        self.toks, self.i = tokens, 0
        self.types = set(BASIC_TYPES)
        self.structs: dict[str, list[tuple[str, str]]] = {}

    def peek(self, k=0):
        j = self.i + k
        return self.toks[j] if j < len(self.toks) else ("eof", "")

    def next(self):
        t = self.peek(); self.i += 1
        return t

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] != "eof":
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            ctx = " ".join(v for _, v in self.toks[max(0, self.i - 8):self.i + 4])
            raise SyntaxError(f"GLSL: expected {val!r}, got {self.peek()[1]!r} near: {ctx}")

    # ---- top level
    def translation_unit(self):
        items = []
        while self.peek()[0] != "eof":
            if self.accept(";"):
                continue
            items.append(self.external())
        return items

    def qualifiers(self):
        q = []
        while self.peek()[1] in QUALIFIERS or self.peek()[1] == "layout":
            if self.next()[1] == "layout":
                self.expect("(")
                while not self.accept(")"):
                    self.next()
            else:
                q.append(self.toks[self.i - 1][1])
        return q

    def type_spec(self):
        if self.peek()[1] == "struct":
            return self.struct_def()
        k, v = self.next()
        if v not in self.types:
            raise SyntaxError(f"GLSL: unknown type {v!r}")
        return v

    def array_suffix(self, t):
        while self.accept("["):
            n = None if self.peek()[1] == "]" else self.expr()
            self.expect("]")
            t = ("array", t, n)
        return t

    def struct_def(self):
        self.expect("struct")
        name = self.next()[1]
        self.expect("{")
        fields = []
        while not self.accept("}"):
            self.qualifiers()
            t = self.type_spec()
            while True:
                fname = self.next()[1]
                fields.append((self.array_suffix(t), fname))
                if not self.accept(","):
                    break
            self.expect(";")
        self.types.add(name)
        self.structs[name] = fields
        return name

    def external(self):
        quals = self.qualifiers()
        if self.peek()[1] == "precision":
            while self.next()[1] != ";":
                pass
            return ("nop",)
        t = self.type_spec()
        if self.accept(";"):                       # bare struct definition
            return ("nop",)
        t = self.array_suffix(t)
        name = self.next()[1]
        if self.accept("("):                       # function
            params = []
            if not self.accept(")"):
                while True:
                    pq = self.qualifiers()
                    pt = self.array_suffix(self.type_spec())
                    pname = None
                    if self.peek()[0] == "id":
                        pname = self.next()[1]
                        pt = self.array_suffix(pt)
                    if not (pt == "void" and pname is None):
                        direction = "inout" if "inout" in pq else "out" if "out" in pq else "in"
                        params.append((direction, pt, pname))
                    if self.accept(")"):
                        break
                    self.expect(",")
            if self.accept(";"):
                return ("nop",)
            body = self.compound()
            return ("function", t, name, params, body)
        decls = self.declarators(t, name)
        self.expect(";")
        return ("global", quals, decls)

    def declarators(self, t, first_name):
        decls, name = [], first_name
        while True:
            dt = self.array_suffix(t)
            init = self.assignment() if self.accept("=") else None
            decls.append((dt, name, init))
            if not self.accept(","):
                return decls
            name = self.next()[1]

    # ---- statements
    def compound(self):
        self.expect("{")
        body = []
        while not self.accept("}"):
            body.append(self.statement())
        return ("block", body)

    def is_declaration(self):
        j = 0
        while self.peek(j)[1] in QUALIFIERS:
            j += 1
        k, v = self.peek(j)
        if v == "struct":
            return True
        if v not in self.types:
            return False
        j += 1
        while self.peek(j)[1] == "[":               # type[n] name  vs  type[n](...)
            depth = 0
            while True:
                v2 = self.peek(j)[1]
                depth += v2 == "["
                depth -= v2 == "]"
                j += 1
                if depth == 0:
                    break
        return self.peek(j)[0] == "id" and self.peek(j)[1] not in self.types

    def statement(self):
        k, v = self.peek()
        if v == "{":
            return self.compound()
        if v == ";":
            self.next()
            return ("nop",)
        if v == "if":
            self.next(); self.expect("(")
            c = self.expr(); self.expect(")")
            a = self.statement()
            b = self.statement() if self.accept("else") else None
            return ("if", c, a, b)
        if v == "for":
            self.next(); self.expect("(")
            init = ("nop",) if self.accept(";") else self.simple_statement()
            cond = None if self.peek()[1] == ";" else self.expr()
            self.expect(";")
            step = None if self.peek()[1] == ")" else self.expr()
            self.expect(")")
            return ("for", init, cond, step, self.statement())
        if v == "while":
            self.next(); self.expect("(")
            c = self.expr(); self.expect(")")
            return ("for", ("nop",), c, None, self.statement())
        if v == "do":
            self.next()
            body = self.statement()
            self.expect("while"); self.expect("(")
            c = self.expr(); self.expect(")"); self.expect(";")
            return ("dowhile", body, c)
        if v == "switch":
            self.next(); self.expect("(")
            sel = self.expr(); self.expect(")"); self.expect("{")
            items = []
            while not self.accept("}"):
                if self.accept("case"):
                    items.append(("case", self.expr())); self.expect(":")
                elif self.accept("default"):
                    self.expect(":"); items.append(("default",))
                else:
                    items.append(self.statement())
            return ("switch", sel, items)
        if v in ("break", "continue", "discard"):
            self.next(); self.expect(";")
            return (v,)
        if v == "return":
            self.next()
            e = None if self.peek()[1] == ";" else self.expr()
            self.expect(";")
            return ("return", e)
        return self.simple_statement()

    def simple_statement(self):
        if self.is_declaration():
            quals = self.qualifiers()
            t = self.type_spec()
            t = self.array_suffix(t)
            name = self.next()[1]
            decls = self.declarators(t, name)
            self.expect(";")
            return ("decl", quals, decls)
        e = self.expr()
        self.expect(";")
        return ("expr", e)

    # ---- expressions
    def expr(self):
        e = self.assignment()
        while self.accept(","):
            e = ("comma", e, self.assignment())
        return e

    ASSIGN = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>="}

    def assignment(self):
        left = self.ternary()
        if self.peek()[1] in self.ASSIGN and self.peek()[0] == "op":
            op = self.next()[1]
            return ("assign", op, left, self.assignment())
        return left

    def ternary(self):
        c = self.binary(0)
        if self.accept("?"):
            a = self.expr()
            self.expect(":")
            return ("ternary", c, a, self.assignment())
        return c

    LEVELS = [["||"], ["^^"], ["&&"], ["|"], ["^"], ["&"], ["==", "!="], ["<", ">", "<=", ">="], ["<<", ">>"],
              ["+", "-"], ["*", "/", "%"]]

    def binary(self, level):
        if level == len(self.LEVELS):
            return self.unary()
        left = self.binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self.LEVELS[level]:
            op = self.next()[1]
            left = ("binary", op, left, self.binary(level + 1))
        return left

    def unary(self):
        k, v = self.peek()
        if k == "op" and v in ("+", "-", "!", "~"):
            self.next()
            return ("unary", v, self.unary())
        if k == "op" and v in ("++", "--"):
            self.next()
            return ("preinc", v, self.unary())
        return self.postfix()

    def postfix(self):
        e = self.primary()
        while True:
            if self.accept("["):
                e = ("index", e, self.expr()); self.expect("]")
            elif self.accept("."):
                name = self.next()[1]
                if name == "length" and self.peek()[1] == "(":
                    self.next(); self.expect(")")
                    e = ("length", e)
                else:
                    e = ("field", e, name)
            elif self.peek()[1] in ("++", "--") and self.peek()[0] == "op":
                e = ("postinc", self.next()[1], e)
            else:
                return e

    def primary(self):
        k, v = self.next()
        if k == "float":
            return ("lit", "float", float(v.rstrip("fF")))
        if k == "int":
            digits = v.rstrip("uU")
            value = int(digits, 0) if digits.lower().startswith("0x") or digits == "0" or not digits.startswith("0") else int(digits, 8)
            return ("lit", "uint" if v[-1] in "uU" else "int", value)
        if v == "(":
            e = self.expr(); self.expect(")")
            return e
        if k == "id":
            if v in ("true", "false"):
                return ("lit", "bool", v == "true")
            if v in self.types:
                t = self.array_suffix(v)
                self.expect("(")
                return ("construct", t, self.args())
            if self.accept("("):
                return ("call", v, self.args())
            return ("name", v)
        raise SyntaxError(f"GLSL: unexpected token {v!r}")

    def args(self):
        a = []
        if self.accept(")"):
            return a
        while True:
            a.append(self.assignment())
            if self.accept(")"):
                return a
            self.expect(",")

# ---------------------------------------------------------------------------------------------- #
# Machine

SWIZZLE = {c: i for s in ("xyzw", "rgba", "stpq") for i, c in enumerate(s)}


@dataclass
class Frame:
    rtype: object
    returned: np.ndarray | None = None
    retval: V | None = None
    flow: list = field(default_factory=list)       # [{kind, brk, cont}]


class GLSLError(RuntimeError):
    pass


def m_and(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return a & b


def m_or(a, b, n):
    """masks of lanes that LEFT (None = nobody)"""
    if a is None:
        return b
    if b is None:
        return a
    return a | b


class Machine:
    """Executes one parsed shader over `lanes` fragments. Inputs (uniforms, varyings, samplers) are given by
    name; `run()` executes main() and returns the `out` variables."""

    def __init__(self, source: str, defines: dict[str, str] | None = None):
        self.parser = Parser(preprocess(source, defines))
        self.unit = self.parser.translation_unit()
        self.structs = self.parser.structs
        self.functions: dict[str, list] = {}
        self.global_decls = []
        for item in self.unit:
            if item[0] == "function":
                self.functions.setdefault(item[2], []).append(item)
            elif item[0] == "global":
                self.global_decls.append(item)
        self.inputs = {}                           # name → (kind, type) for uniform / in
        self.outputs = {}
        for _, quals, decls in self.global_decls:
            for t, name, init in decls:
                if "uniform" in quals:
                    self.inputs[name] = ("uniform", t)
                elif "in" in quals or "attribute" in quals:
                    self.inputs[name] = ("in", t)
                elif "out" in quals or "varying" in quals:
                    self.outputs[name] = t

    # ---- values
    def const(self, t, value):
        info = vec_info(t)
        if info:
            base, n = info
            arr = np.asarray(value, SCALARS[base])
            if n == 1:
                arr = arr.reshape(-1) if arr.ndim else arr.reshape(1)
            elif arr.ndim == 1:
                arr = arr.reshape(1, n)
            return V(t, arr)
        if t in ("mat2", "mat3", "mat4"):
            n = int(t[3])
            arr = np.asarray(value, F32)
            return V(t, arr.reshape((-1, n, n)) if arr.ndim != 2 else arr.reshape(1, n, n))
        raise GLSLError(f"cannot build a constant of type {t}")

    def zero(self, t):
        if isinstance(t, tuple):
            n = self.array_len(t)
            return V(("array", t[1], n), [self.zero(t[1]) for _ in range(n)])
        info = vec_info(t)
        if info:
            base, n = info
            return V(t, np.zeros((1,) if n == 1 else (1, n), SCALARS[base]))
        if t in ("mat2", "mat3", "mat4"):
            n = int(t[3])
            return V(t, np.zeros((1, n, n), F32))
        if t in self.structs:
            return V(t, {name: self.zero(ft) for ft, name in self.structs[t]})
        if t == "sampler2D":
            return V(t, None)
        raise GLSLError(f"cannot zero-initialise {t}")

    def array_len(self, t):
        n = t[2]
        if isinstance(n, int):
            return n
        v = self.eval(n, None)
        return int(v.a.reshape(-1)[0])

    def convert(self, v: V, t) -> V:
        """implicit conversion on initialisation / assignment / argument passing"""
        if isinstance(t, tuple):
            if isinstance(v.t, tuple):
                return V(("array", t[1], len(v.a)), [self.convert(e, t[1]) for e in v.a])
            raise GLSLError(f"cannot convert {v.t} to array")
        if v.t == t:
            return v
        a, b = vec_info(v.t), vec_info(t)
        if a and b and a[1] == b[1]:
            return V(t, self.cast(v.a, a[0], b[0]))
        raise GLSLError(f"cannot convert {v.t} to {t}")

    @staticmethod
    def cast(arr, src, dst):
        if src == dst:
            return arr
        if dst in ("int", "uint") and src == "float":
            with np.errstate(all="ignore"):
                tr = np.trunc(arr.astype(np.float64))
                tr = np.where(np.isfinite(tr), tr, 0.0)
                return np.clip(tr, -2**31, 2**32 - 1).astype(np.int64).astype(SCALARS[dst])
        if dst == "bool":
            return arr != 0
        return arr.astype(SCALARS[dst])

    def blend(self, old: V | None, new: V, mask) -> V:
        """masked store: lanes in `mask` take `new`, the others keep `old`"""
        if mask is None or old is None:
            return new
        if isinstance(new.a, dict):
            return V(new.t, {k: self.blend(old.a.get(k), new.a[k], mask) for k in new.a})
        if isinstance(new.a, list):
            return V(new.t, [self.blend(o, n, mask) for o, n in zip(old.a, new.a)])
        if new.t == "sampler2D":
            return new
        m = mask.reshape(mask.shape + (1,)*(new.a.ndim - 1))
        return V(new.t, np.where(m, new.a, old.a))

    # ---- scopes
    def lookup(self, name):
        for scope in reversed(self.scopes):
            if name in scope:
                return scope
        raise GLSLError(f"undeclared identifier {name!r}")

    def live(self, mask):
        """`mask` minus the lanes that returned / broke / continued in the current function"""
        fr = self.frame
        left = fr.returned
        for fl in fr.flow:
            left = m_or(left, fl["brk"], self.lanes)
            left = m_or(left, fl["cont"], self.lanes)
        if left is None:
            return mask
        return ~left if mask is None else (mask & ~left)

    def full(self, mask):
        return np.ones(self.lanes, bool) if mask is None else mask

    # ---- execution
    def run(self, lanes: int, inputs: dict, samplers: dict | None = None) -> dict[str, V]:
        """inputs: name → python scalar / sequence / ndarray (lanes, ...) ; samplers: name → Texture"""
        self.lanes = lanes
        self.discarded = None
        g = {}
        self.scopes = [g]
        self.frame = Frame("void")
        with np.errstate(all="ignore"):
            for _, quals, decls in self.global_decls:
                for t, name, init in decls:
                    if name in self.inputs:
                        if t == "sampler2D":
                            tex = (samplers or {}).get(name)
                            g[name] = V(t, tex)
                        elif name in inputs:
                            g[name] = self.input_value(t, inputs[name])
                        else:
                            g[name] = self.zero(t)
                    elif init is not None:
                        g[name] = self.convert(self.eval(init, None), self.resolve_type(t, init))
                    else:
                        g[name] = self.zero(t)
            main = self.functions["main"][0]
            self.call_user(main, [], None)
        return {name: g[name] for name in self.outputs}

    def resolve_type(self, t, init=None):
        if isinstance(t, tuple) and t[2] is None and init is not None and init[0] == "construct":
            return ("array", t[1], len(init[2]))
        return t

    def input_value(self, t, value):
        info = vec_info(t)
        if info is None:
            return self.const(t, value)
        base, n = info
        arr = np.asarray(value)
        if arr.dtype == object:
            raise GLSLError("bad input")
        arr = arr.astype(SCALARS[base])
        if n == 1:
            return V(t, arr.reshape(-1) if arr.ndim else arr.reshape(1))
        if arr.ndim == 1:
            arr = arr.reshape(1, n)
        return V(t, arr.reshape(-1, n))

    def exec(self, node, mask):
        kind = node[0]
        if kind == "block":
            self.scopes.append({})
            try:
                for st in node[1]:
                    m = self.live(mask)
                    if m is not None and not m.any():
                        break
                    self.exec(st, m)
            finally:
                self.scopes.pop()
        elif kind == "expr":
            self.eval(node[1], mask)
        elif kind == "decl":
            for t, name, init in node[2]:
                if init is not None:
                    val = self.eval(init, mask)                 # the name is not in scope in its own initialiser
                    val = self.convert(val, self.resolve_type(t, init))
                else:
                    val = self.zero(t)
                self.scopes[-1][name] = val
        elif kind == "if":
            c = self.eval(node[1], mask)
            if c.a.shape[0] == 1:
                branch = node[2] if bool(c.a[0]) else node[3]
                if branch is not None:
                    self.exec_scoped(branch, mask)
            else:
                mt = m_and(mask, c.a)
                if mt.any():
                    self.exec_scoped(node[2], mt)
                if node[3] is not None:
                    me = m_and(mask, ~c.a)
                    me = self.live(me)
                    if me.any():
                        self.exec_scoped(node[3], me)
        elif kind == "for":
            self.exec_loop(node[1], node[2], node[3], node[4], mask, test_first=True)
        elif kind == "dowhile":
            self.exec_loop(("nop",), node[2], None, node[1], mask, test_first=False)
        elif kind == "switch":
            self.exec_switch(node, mask)
        elif kind == "return":
            fr = self.frame
            if node[1] is not None:
                val = self.convert(self.eval(node[1], mask), fr.rtype)
                fr.retval = self.blend(fr.retval, val, mask) if fr.retval is not None else val
            fr.returned = m_or(fr.returned, self.full(mask), self.lanes)
        elif kind == "break":
            for fl in reversed(self.frame.flow):
                fl["brk"] = m_or(fl["brk"], self.full(mask), self.lanes)
                break
        elif kind == "continue":
            for fl in reversed(self.frame.flow):
                if fl["kind"] == "loop":
                    fl["cont"] = m_or(fl["cont"], self.full(mask), self.lanes)
                    break
        elif kind == "discard":
            self.discarded = m_or(self.discarded, self.full(mask), self.lanes)
            self.frame.returned = m_or(self.frame.returned, self.full(mask), self.lanes)
        elif kind == "nop":
            pass
        else:
            raise GLSLError(f"statement {kind}")

    def exec_scoped(self, node, mask):
        if node[0] == "block":
            self.exec(node, mask)
        else:
            self.scopes.append({})
            try:
                self.exec(node, mask)
            finally:
                self.scopes.pop()

    def exec_loop(self, init, cond, step, body, mask, test_first):
        self.scopes.append({})
        fl = {"kind": "loop", "brk": None, "cont": None}
        self.frame.flow.append(fl)
        try:
            self.exec(init, mask)
            lm, first = mask, True
            for _ in range(1 << 20):
                if cond is not None and (test_first or not first):
                    c = self.eval(cond, lm)
                    if c.a.shape[0] == 1:
                        if not bool(c.a[0]):
                            break
                    else:
                        lm = m_and(lm, c.a)
                first = False
                lm = self.live(lm)
                if lm is not None and not lm.any():
                    break
                self.exec_scoped(body, lm)
                fl["cont"] = None
                lm = self.live(lm)
                if lm is not None and not lm.any():
                    break
                if step is not None:
                    self.eval(step, lm)
            else:
                raise GLSLError("loop did not terminate")
        finally:
            self.frame.flow.pop()
            self.scopes.pop()

    def exec_switch(self, node, mask):
        sel = self.eval(node[1], mask)
        labels = [self.eval(it[1], None) for it in node[2] if it[0] == "case"]
        matched_any = None
        for lab in labels:
            matched_any = m_or(matched_any, np.broadcast_to(sel.a == lab.a, (self.lanes,)), self.lanes)
        fl = {"kind": "switch", "brk": None, "cont": None}
        self.frame.flow.append(fl)
        self.scopes.append({})
        try:
            entered = np.zeros(self.lanes, bool)
            for it in node[2]:
                if it[0] == "case":
                    lab = self.eval(it[1], None)
                    entered = entered | m_and(self.full(mask), np.broadcast_to(sel.a == lab.a, (self.lanes,)))
                elif it[0] == "default":
                    none = self.full(mask) if matched_any is None else (self.full(mask) & ~matched_any)
                    entered = entered | none
                else:
                    m = self.live(entered)
                    if m.any():
                        self.exec(it, m)
        finally:
            self.scopes.pop()
            self.frame.flow.pop()

    # ---- expressions
    def eval(self, node, mask) -> V:
        kind = node[0]
        if kind == "lit":
            return V(node[1], np.asarray([node[2]], SCALARS[node[1]]))
        if kind == "name":
            return self.lookup(node[1])[node[1]]
        if kind == "binary":
            op = node[1]
            if op in ("&&", "||"):
                a = self.eval(node[2], mask)
                if a.a.shape[0] == 1:
                    if bool(a.a[0]) == (op == "||"):
                        return a
                    return self.eval(node[3], mask)
                sub = m_and(mask, a.a if op == "&&" else ~a.a)
                b = self.eval(node[3], sub)
                return V("bool", (a.a & b.a) if op == "&&" else (a.a | b.a))
            return self.binary(op, self.eval(node[2], mask), self.eval(node[3], mask))
        if kind == "unary":
            v = self.eval(node[2], mask)
            if node[1] == "-":
                return V(v.t, -v.a)
            if node[1] == "!":
                return V(v.t, ~v.a)
            if node[1] == "~":
                return V(v.t, ~v.a)
            return v
        if kind == "assign":
            op, target = node[1], node[2]
            val = self.eval(node[3], mask)
            if op != "=":
                val = self.binary(op[:-1], self.eval(target, mask), val)
            return self.assign(target, val, mask)
        if kind in ("preinc", "postinc"):
            old = self.eval(node[2], mask)
            one = V("int", np.asarray([1], I32))
            new = self.binary("+" if node[1] == "++" else "-", old, one)
            self.assign(node[2], new, mask)
            return new if kind == "preinc" else old
        if kind == "ternary":
            c = self.eval(node[1], mask)
            if c.a.shape[0] == 1:
                return self.eval(node[2] if bool(c.a[0]) else node[3], mask)
            a = self.eval(node[2], m_and(mask, c.a))
            b = self.eval(node[3], m_and(mask, ~c.a))
            a, b = self.unify(a, b)
            return self.blend(b, a, c.a)
        if kind == "field":
            return self.field(self.eval(node[1], mask), node[2])
        if kind == "index":
            return self.index(self.eval(node[1], mask), self.eval(node[2], mask))
        if kind == "length":                       # .length() of an array, a vector (components) or a matrix (columns)
            v = self.eval(node[1], mask)
            n = len(v.a) if isinstance(v.a, list) else v.a.shape[1]
            return V("int", np.asarray([n], I32))
        if kind == "construct":
            return self.construct(node[1], [self.eval(a, mask) for a in node[2]])
        if kind == "call":
            return self.call(node[1], node[2], mask)
        if kind == "comma":
            self.eval(node[1], mask)
            return self.eval(node[2], mask)
        raise GLSLError(f"expression {kind}")

    def unify(self, a: V, b: V):
        if a.t == b.t:
            return a, b
        ia, ib = vec_info(a.t), vec_info(b.t)
        if ia and ib and ia[1] == ib[1]:
            base = "float" if "float" in (ia[0], ib[0]) else "uint" if "uint" in (ia[0], ib[0]) else ia[0]
            t = make_type(base, ia[1])
            return self.convert(a, t), self.convert(b, t)
        raise GLSLError(f"type mismatch {a.t} / {b.t}")

    def field(self, v: V, name: str) -> V:
        if isinstance(v.a, dict):
            return v.a[name]
        if isinstance(v.t, tuple) and name == "length":
            raise GLSLError(".length() unsupported")
        info = vec_info(v.t)
        if info and info[1] > 1 and all(c in SWIZZLE for c in name):
            idx = [SWIZZLE[c] for c in name]
            if len(idx) == 1:
                return V(info[0], v.a[:, idx[0]])
            return V(make_type(info[0], len(idx)), v.a[:, idx])
        raise GLSLError(f"no field {name!r} on {v.t}")

    def index(self, v: V, i: V) -> V:
        uniform = i.a.shape[0] == 1
        if isinstance(v.a, list):
            if uniform:
                return v.a[int(i.a[0])]
            out = v.a[0]
            for k in range(1, len(v.a)):
                out = self.blend(out, v.a[k], i.a == k)
            return out
        info = vec_info(v.t)
        if info and info[1] > 1:
            if uniform:
                return V(info[0], v.a[:, int(i.a[0])])
            arr = np.broadcast_to(v.a, (self.lanes, info[1]))
            return V(info[0], np.take_along_axis(arr, np.clip(i.a, 0, info[1] - 1)[:, None].astype(np.int64), 1)[:, 0])
        if v.t in ("mat2", "mat3", "mat4") and uniform:
            return V("vec" + v.t[3], v.a[:, int(i.a[0]), :])
        raise GLSLError(f"cannot index {v.t}")

    # ---- assignment through lvalues
    def assign(self, target, val: V, mask) -> V:
        kind = target[0]
        if kind == "name":
            scope = self.lookup(target[1])
            old = scope[target[1]]
            val = self.convert(val, old.t)
            scope[target[1]] = self.blend(old, val, mask)
            return val
        if kind == "field":
            base = self.eval(target[1], mask)
            name = target[2]
            if isinstance(base.a, dict):
                val = self.convert(val, base.a[name].t)
                new = dict(base.a)
                new[name] = val
                self.assign(target[1], V(base.t, new), mask)
                return val
            info = vec_info(base.t)
            idx = [SWIZZLE[c] for c in name]
            val = self.convert(val, make_type(info[0], len(idx)))
            lanes = max(base.a.shape[0], val.a.shape[0])
            arr = np.broadcast_to(base.a, (lanes, info[1])).copy()
            arr[:, idx] = val.a.reshape(val.a.shape[0], len(idx))
            self.assign(target[1], V(base.t, arr), mask)
            return val
        if kind == "index":
            base = self.eval(target[1], mask)
            i = self.eval(target[2], mask)
            if i.a.shape[0] != 1:
                raise GLSLError("store through a divergent index is not supported")
            k = int(i.a[0])
            if isinstance(base.a, list):
                new = list(base.a)
                new[k] = self.convert(val, base.t[1])
                self.assign(target[1], V(base.t, new), mask)
                return val
            info = vec_info(base.t)
            if info and info[1] > 1:
                val = self.convert(val, info[0])
                lanes = max(base.a.shape[0], val.a.shape[0])
                arr = np.broadcast_to(base.a, (lanes, info[1])).copy()
                arr[:, k] = val.a
                self.assign(target[1], V(base.t, arr), mask)
                return val
            if isinstance(base.t, str) and base.t.startswith("mat"):          # m[k] = column
                val = self.convert(val, "vec" + base.t[3])
                lanes = max(base.a.shape[0], val.a.shape[0])
                arr = np.broadcast_to(base.a, (lanes,) + base.a.shape[1:]).copy()
                arr[:, k, :] = val.a
                self.assign(target[1], V(base.t, arr), mask)
                return val
        raise GLSLError(f"not an lvalue: {kind}")

    # ---- operators
    def binary(self, op, a: V, b: V) -> V:
        if op in ("==", "!="):
            eq = self.equal(a, b)
            return V("bool", eq if op == "==" else ~eq)
        ta, tb = a.t, b.t
        mats = ("mat2", "mat3", "mat4")
        if ta in mats or tb in mats:
            return self.matrix_op(op, a, b)
        ia, ib = vec_info(ta), vec_info(tb)
        if ia is None or ib is None:
            raise GLSLError(f"operator {op} on {ta}, {tb}")
        if op in ("<", ">", "<=", ">="):
            x, y, _ = self.promote(a, b, ia, ib)
            fn = {"<": np.less, ">": np.greater, "<=": np.less_equal, ">=": np.greater_equal}[op]
            return V("bool", fn(x, y))
        if op == "^^":
            return V("bool", a.a ^ b.a)
        x, y, base = self.promote(a, b, ia, ib)
        n = max(ia[1], ib[1])
        if ia[1] != ib[1]:
            if ia[1] == 1:
                x = x[:, None]
            elif ib[1] == 1:
                y = y[:, None]
            else:
                raise GLSLError(f"operator {op} on {ta}, {tb}")
        t = make_type(base, n)
        if base == "float":
            if op == "+": r = x + y
            elif op == "-": r = x - y
            elif op == "*": r = x * y
            elif op == "/": r = x / y
            else: raise GLSLError(f"operator {op} on floats")
            return V(t, r.astype(F32, copy=False))
        if op == "+": r = x + y
        elif op == "-": r = x - y
        elif op == "*": r = x * y
        elif op == "/":
            r = np.fix(x.astype(np.float64)/np.where(y == 0, 1, y).astype(np.float64)).astype(x.dtype)
        elif op == "%":
            r = np.fmod(x, np.where(y == 0, 1, y))
        elif op == "&": r = x & y
        elif op == "|": r = x | y
        elif op == "^": r = x ^ y
        elif op == "<<": r = x << y
        elif op == ">>": r = x >> y
        else: raise GLSLError(f"operator {op}")
        return V(t, r.astype(SCALARS[base], copy=False))

    def promote(self, a, b, ia, ib):
        if ia[0] == ib[0]:
            return a.a, b.a, ia[0]
        if "float" in (ia[0], ib[0]):
            return self.cast(a.a, ia[0], "float"), self.cast(b.a, ib[0], "float"), "float"
        if "uint" in (ia[0], ib[0]):
            return self.cast(a.a, ia[0], "uint"), self.cast(b.a, ib[0], "uint"), "uint"
        raise GLSLError(f"no implicit conversion between {a.t} and {b.t}")

    def equal(self, a: V, b: V):
        if isinstance(a.a, dict):
            out = None
            for k in a.a:
                e = self.equal(a.a[k], b.a[k])
                out = e if out is None else out & e
            return out
        if a.t in ("mat2", "mat3", "mat4") and a.t == b.t:          # §5.9: equal when every component is
            e = np.broadcast_arrays(a.a, b.a)[0] == np.broadcast_arrays(a.a, b.a)[1]
            return e.all(axis=(1, 2))
        ia, ib = vec_info(a.t), vec_info(b.t)
        x, y, _ = self.promote(a, b, ia, ib)
        e = x == y
        return e if e.ndim == 1 else e.all(axis=tuple(range(1, e.ndim)))

    def matrix_op(self, op, a: V, b: V) -> V:
        mats = ("mat2", "mat3", "mat4")
        if a.t in mats and b.t in mats:
            if op == "*":
                n = int(a.t[3])
                cols = []
                for c in range(n):                   # result column c = A · B[c]
                    cols.append(self.mat_vec(a.a, b.a[:, c, :], n))
                return V(a.t, np.stack(cols, 1))
            return V(a.t, {"+": np.add, "-": np.subtract, "/": np.divide}[op](a.a, b.a).astype(F32))
        if a.t in mats:
            ib = vec_info(b.t)
            n = int(a.t[3])
            y = self.cast(b.a, ib[0], "float")
            if ib[1] == 1:
                return V(a.t, {"*": np.multiply, "/": np.divide, "+": np.add, "-": np.subtract}[op](a.a, y[:, None, None]).astype(F32))
            if op != "*" or ib[1] != n:
                raise GLSLError("matrix · vector shape")
            return V(b.t if ib[0] == "float" else "vec" + str(n), self.mat_vec(a.a, y, n))
        ia = vec_info(a.t)
        n = int(b.t[3])
        x = self.cast(a.a, ia[0], "float")
        if ia[1] == 1:
            return V(b.t, {"*": np.multiply, "/": np.divide, "+": np.add, "-": np.subtract}[op](x[:, None, None], b.a).astype(F32))
        if op != "*" or ia[1] != n:
            raise GLSLError("vector · matrix shape")
        out = []
        for c in range(n):                           # (v·M)[c] = dot(v, M[c])
            acc = x[:, 0]*b.a[:, c, 0]
            for r in range(1, n):
                acc = acc + x[:, r]*b.a[:, c, r]
            out.append(acc)
        return V("vec" + str(n), np.stack(np.broadcast_arrays(*out), 1).astype(F32))

    @staticmethod
    def mat_vec(m, v, n):
        """(M·v)[r] = Σ_c M[c][r]·v[c], summed left to right (column-major matrices)"""
        rows = []
        for r in range(n):
            acc = m[:, 0, r]*v[:, 0]
            for c in range(1, n):
                acc = acc + m[:, c, r]*v[:, c]
            rows.append(acc)
        return np.stack(np.broadcast_arrays(*rows), 1).astype(F32)

    # ---- constructors
    def construct(self, t, args: list[V]) -> V:
        if isinstance(t, tuple):
            n = len(args) if t[2] is None else self.array_len(t)
            if len(args) != n:
                raise GLSLError("array constructor arity")
            return V(("array", t[1], n), [self.convert(a, t[1]) for a in args])
        if t in self.structs:
            fields = self.structs[t]
            return V(t, {name: self.convert(a, ft) for (ft, name), a in zip(fields, args)})
        info = vec_info(t)
        if info:
            base, n = info
            parts = []
            for a in args:
                if a.t in ("mat2", "mat3", "mat4"):
                    arr = a.a.reshape(a.a.shape[0], -1)
                    src = "float"
                else:
                    ai = vec_info(a.t)
                    if ai is None:
                        raise GLSLError(f"cannot construct {t} from {a.t}")
                    arr, src = (a.a[:, None] if ai[1] == 1 else a.a), ai[0]
                parts.append(self.cast(arr, src, base))
            if n == 1:
                return V(t, parts[0][:, 0])
            total = sum(p.shape[1] for p in parts)
            if len(parts) == 1 and total == 1:
                lanes = parts[0].shape[0]
                return V(t, np.broadcast_to(parts[0], (lanes, n)).copy())
            if total < n:
                raise GLSLError(f"too few components for {t}")
            lanes = max(p.shape[0] for p in parts)
            cat = np.concatenate([np.broadcast_to(p, (lanes, p.shape[1])) for p in parts], 1)
            return V(t, np.ascontiguousarray(cat[:, :n]))
        if t in ("mat2", "mat3", "mat4"):
            n = int(t[3])
            if len(args) == 1 and vec_info(args[0].t) and vec_info(args[0].t)[1] == 1:
                s = self.cast(args[0].a, vec_info(args[0].t)[0], "float")
                m = np.zeros((s.shape[0], n, n), F32)
                for k in range(n):
                    m[:, k, k] = s
                return V(t, m)
            if len(args) == 1 and args[0].t in ("mat2", "mat3", "mat4"):
                k = int(args[0].t[3])
                m = np.zeros((args[0].a.shape[0], n, n), F32)
                for d in range(n):
                    m[:, d, d] = 1
                lo = min(n, k)
                m[:, :lo, :lo] = args[0].a[:, :lo, :lo]
                return V(t, m)
            flat = self.construct("vec4", args[:1]) if False else None
            parts = []
            for a in args:
                ai = vec_info(a.t)
                arr = a.a[:, None] if ai[1] == 1 else a.a
                parts.append(self.cast(arr, ai[0], "float"))
            lanes = max(p.shape[0] for p in parts)
            cat = np.concatenate([np.broadcast_to(p, (lanes, p.shape[1])) for p in parts], 1)
            if cat.shape[1] != n*n:
                raise GLSLError(f"{t} constructor needs {n*n} components")
            return V(t, cat.reshape(lanes, n, n).astype(F32))     # column-major: consecutive values fill a column
        raise GLSLError(f"constructor {t}")

    # ---- calls
    def call(self, name, arg_nodes, mask) -> V:
        if name in self.functions:
            args = [self.eval(a, mask) for a in arg_nodes]
            fn = self.resolve(name, args)
            return self.call_user(fn, args, mask, arg_nodes)
        args = [self.eval(a, mask) for a in arg_nodes]
        if name in ("modf", "frexp"):                      # the two builtins with an out parameter (§8.3)
            result, second = self.builtin(name, args[:1])
            self.assign(arg_nodes[1], second, mask)
            return result
        return self.builtin(name, args)

    def resolve(self, name, args):
        cands = [f for f in self.functions[name] if len(f[3]) == len(args)]
        for f in cands:
            if all(p[1] == a.t for p, a in zip(f[3], args)):
                return f
        def convertible(a, t):
            if not isinstance(a.t, str) or not isinstance(t, str):          # arrays: same element type
                return (not isinstance(a.t, str)) and (not isinstance(t, str)) and a.t[1] == t[1]
            ia, ib = vec_info(a.t), (vec_info(t) if isinstance(t, str) else None)
            return a.t == t or (ia and ib and ia[1] == ib[1] and ia[0] in ("int", "uint") and ib[0] == "float") \
                or (ia and ib and ia[1] == ib[1] and ia[0] == "int" and ib[0] == "uint")
        ok = [f for f in cands if all(convertible(a, p[1]) for p, a in zip(f[3], args))]
        if not ok:
            raise GLSLError(f"no overload of {name} for ({', '.join(str(a.t) for a in args)})")
        return ok[0]

    def call_user(self, fn, args, mask, arg_nodes=None) -> V:
        _, rtype, name, params, body = fn
        scope = {}
        for (direction, pt, pname), a in zip(params, args):
            if pname is None:
                continue
            scope[pname] = self.zero(pt) if direction == "out" else self.convert(a, pt)
        saved_scopes, saved_frame = self.scopes, self.frame
        self.scopes = [saved_scopes[0], scope]
        self.frame = Frame(rtype)
        try:
            self.exec(body, mask)
            result = self.frame.retval
        finally:
            self.scopes, self.frame = saved_scopes, saved_frame
        for k, (direction, pt, pname) in enumerate(params):
            if direction in ("out", "inout") and arg_nodes is not None:
                self.assign(arg_nodes[k], scope[pname], mask)
        if rtype == "void":
            return V("void", None)
        if result is None:
            return self.zero(rtype)
        return result

    # ---- builtins (GLSL 3.30 §8)
    def builtin(self, name, args: list[V]) -> V:
        fn = getattr(self, "b_" + name, None)
        if fn is None:
            raise GLSLError(f"unknown function {name!r}")
        return fn(*args)

    @staticmethod
    def fl(v: V):
        """→ (float array, type as float vector, dim)"""
        base, n = vec_info(v.t)
        return Machine.cast(v.a, base, "float"), make_type("float", n), n

    def gen(self, *vs):
        """genType broadcasting: scalars stretch over the widest vector argument"""
        arrs, n = [], 1
        for v in vs:
            a, _, k = self.fl(v)
            arrs.append((a, k)); n = max(n, k)
        out = [a[:, None] if (k == 1 and n > 1) else a for a, k in arrs]
        return out, make_type("float", n)

    def _map(self, fn, v):
        a, t, _ = self.fl(v)
        return V(t, fn(a).astype(F32, copy=False))

    def b_radians(self, x): return self._map(lambda a: a*F32(np.pi)/F32(180), x)
    def b_degrees(self, x): return self._map(lambda a: a*F32(180)/F32(np.pi), x)
    def b_sin(self, x): return self._map(np.sin, x)
    def b_cos(self, x): return self._map(np.cos, x)
    def b_tan(self, x): return self._map(np.tan, x)
    def b_asin(self, x): return self._map(np.arcsin, x)
    def b_acos(self, x): return self._map(np.arccos, x)
    def b_sinh(self, x): return self._map(np.sinh, x)
    def b_cosh(self, x): return self._map(np.cosh, x)
    def b_tanh(self, x): return self._map(np.tanh, x)
    def b_exp(self, x): return self._map(np.exp, x)
    def b_log(self, x): return self._map(np.log, x)
    def b_exp2(self, x): return self._map(np.exp2, x)
    def b_log2(self, x): return self._map(np.log2, x)
    def b_sqrt(self, x): return self._map(np.sqrt, x)
    def b_inversesqrt(self, x): return self._map(lambda a: F32(1)/np.sqrt(a), x)
    def b_floor(self, x): return self._map(np.floor, x)
    def b_ceil(self, x): return self._map(np.ceil, x)
    def b_trunc(self, x): return self._map(np.trunc, x)
    def b_round(self, x): return self._map(np.rint, x)
    def b_roundEven(self, x): return self._map(np.rint, x)
    def b_fract(self, x): return self._map(lambda a: a - np.floor(a), x)

    def b_atan(self, y, x=None):
        if x is None:
            return self._map(np.arctan, y)
        (a, b), t = self.gen(y, x)
        return V(t, np.arctan2(a, b).astype(F32))

    def b_pow(self, x, y):
        (a, b), t = self.gen(x, y)
        return V(t, np.power(a, b).astype(F32))

    def b_abs(self, x):
        base, n = vec_info(x.t)
        return V(x.t, np.abs(x.a))

    def b_sign(self, x):
        return V(x.t, np.sign(x.a).astype(x.a.dtype))

    def b_mod(self, x, y):
        (a, b), t = self.gen(x, y)
        return V(t, (a - b*np.floor(a/b)).astype(F32))

    def _minmax(self, fn, x, y):
        ix, iy = vec_info(x.t), vec_info(y.t)
        if ix[0] != "float" and iy[0] != "float":
            a, b = x.a, y.a
            if ix[1] != iy[1]:
                a, b = (a[:, None], b) if ix[1] == 1 else (a, b[:, None])
            return V(make_type(ix[0], max(ix[1], iy[1])), fn(a, b))
        (a, b), t = self.gen(x, y)
        return V(t, fn(a, b).astype(F32))

    def b_min(self, x, y): return self._minmax(np.fmin, x, y)
    def b_max(self, x, y): return self._minmax(np.fmax, x, y)

    def b_clamp(self, x, lo, hi):
        return self.b_min(self.b_max(x, lo), hi)

    def b_mix(self, x, y, a):
        if vec_info(a.t)[0] == "bool":
            (p, q), t = self.gen(x, y)
            sel = a.a if a.a.ndim == p.ndim else a.a[:, None]
            return V(t, np.where(sel, q, p).astype(F32))
        (p, q, w), t = self.gen(x, y, a)
        return V(t, (p*(F32(1) - w) + q*w).astype(F32))

    def b_step(self, edge, x):
        (e, a), t = self.gen(edge, x)
        return V(t, np.where(a < e, F32(0), F32(1)).astype(F32))

    def b_smoothstep(self, e0, e1, x):
        (a, b, v), t = self.gen(e0, e1, x)
        u = np.fmin(np.fmax((v - a)/(b - a), F32(0)), F32(1)).astype(F32)
        return V(t, (u*u*(F32(3) - F32(2)*u)).astype(F32))

    @staticmethod
    def _dot(a, b):
        if a.ndim == 1:
            return (a*b).astype(F32)
        acc = a[:, 0]*b[:, 0]
        for k in range(1, max(a.shape[1], b.shape[1])):
            acc = acc + a[:, k]*b[:, k]
        return acc.astype(F32)

    def b_dot(self, x, y):
        (a, b), _ = self.gen(x, y)
        return V("float", self._dot(a, b))

    def b_length(self, x):
        a, _, n = self.fl(x)
        return V("float", np.abs(a) if n == 1 else np.sqrt(self._dot(a, a)))

    def b_distance(self, x, y):
        return self.b_length(self.binary("-", x, y))

    def b_normalize(self, x):
        a, t, n = self.fl(x)
        if n == 1:
            return V(t, np.sign(a))
        return V(t, (a/np.sqrt(self._dot(a, a))[:, None]).astype(F32))

    def b_cross(self, x, y):
        (a, b), t = self.gen(x, y)
        c = [a[:, 1]*b[:, 2] - b[:, 1]*a[:, 2], a[:, 2]*b[:, 0] - b[:, 2]*a[:, 0], a[:, 0]*b[:, 1] - b[:, 0]*a[:, 1]]
        return V("vec3", np.stack(np.broadcast_arrays(*c), 1).astype(F32))

    def b_reflect(self, i, n):
        (a, b), t = self.gen(i, n)
        d = self._dot(b, a)
        return V(t, (a - F32(2)*d[:, None]*b).astype(F32))

    def b_refract(self, i, n, eta):                        # GLSL 3.30 §8.4: k = 1 − η²(1 − (n·i)²); 0 when k < 0
        (a, b), t = self.gen(i, n)
        e = self.fl(eta)[0].astype(F32)
        d = self._dot(b, a)
        k = (F32(1) - e*e*(F32(1) - d*d).astype(F32)).astype(F32)
        with np.errstate(invalid="ignore"):
            bent = ((e[:, None]*a).astype(F32) - ((e*d).astype(F32) + np.sqrt(k).astype(F32)).astype(F32)[:, None]*b).astype(F32)
        return V(t, np.where((k < 0)[:, None], F32(0), bent).astype(F32))

    def b_faceforward(self, n, i, nref):
        (a, b, c), t = self.gen(n, i, nref)
        return V(t, np.where((self._dot(c, b) < 0)[:, None], a, -a).astype(F32))

    def _bits(self, x, src, dst, base):
        _, n = vec_info(x.t)
        return V(make_type(base, n), np.ascontiguousarray(x.a.astype(src)).view(dst))

    def b_floatBitsToInt(self, x): return self._bits(x, F32, I32, "int")
    def b_floatBitsToUint(self, x): return self._bits(x, F32, U32, "uint")
    def b_intBitsToFloat(self, x): return self._bits(x, I32, F32, "float")
    def b_uintBitsToFloat(self, x): return self._bits(x, U32, F32, "float")

    def b_outerProduct(self, c, r):                        # column j = c·r[j]; square results only (mat2 / mat3 / mat4)
        a, b = self.fl(c)[0], self.fl(r)[0]
        if a.shape[1] != b.shape[1]:
            raise GLSLError("outerProduct of vectors of different sizes (non-square matrices are not modelled)")
        return V(f"mat{a.shape[1]}", (b[:, :, None]*a[:, None, :]).astype(F32))

    def b_matrixCompMult(self, x, y): return V(x.t, (x.a*y.a).astype(F32))

    def b_isnan(self, x): return V(make_type("bool", vec_info(x.t)[1]), np.isnan(x.a))
    def b_isinf(self, x): return V(make_type("bool", vec_info(x.t)[1]), np.isinf(x.a))
    def b_any(self, x): return V("bool", x.a.any(axis=1))
    def b_all(self, x): return V("bool", x.a.all(axis=1))
    def b_not(self, x): return V(x.t, ~x.a)

    def _compare(self, fn, x, y):
        ix, iy = vec_info(x.t), vec_info(y.t)
        a, b, _ = self.promote(x, y, ix, iy)
        return V(make_type("bool", ix[1]), fn(a, b))

    def b_lessThan(self, x, y): return self._compare(np.less, x, y)
    def b_lessThanEqual(self, x, y): return self._compare(np.less_equal, x, y)
    def b_greaterThan(self, x, y): return self._compare(np.greater, x, y)
    def b_greaterThanEqual(self, x, y): return self._compare(np.greater_equal, x, y)
    def b_equal(self, x, y): return self._compare(np.equal, x, y)
    def b_notEqual(self, x, y): return self._compare(np.not_equal, x, y)

    def b_transpose(self, m): return V(m.t, np.swapaxes(m.a, 1, 2).copy())

    def b_determinant(self, m):
        a = m.a.astype(F32)
        if m.t == "mat2":
            return V("float", (a[:, 0, 0]*a[:, 1, 1] - a[:, 1, 0]*a[:, 0, 1]).astype(F32))
        return V("float", np.linalg.det(np.swapaxes(a, 1, 2).astype(np.float64)).astype(F32))

    def b_inverse(self, m):
        a = m.a.astype(F32)
        if m.t == "mat2":                                  # adjugate / determinant, one float32 rounding per operation
            d = (F32(1)/self.b_determinant(m).a).astype(F32)
            out = np.empty_like(a)
            out[:, 0, 0], out[:, 0, 1] = a[:, 1, 1]*d, -a[:, 0, 1]*d
            out[:, 1, 0], out[:, 1, 1] = -a[:, 1, 0]*d, a[:, 0, 0]*d
            return V(m.t, out.astype(F32))
        inv = np.linalg.inv(np.swapaxes(a, 1, 2).astype(np.float64))
        return V(m.t, np.swapaxes(inv, 1, 2).astype(F32))

    # texture access: OpenGL fixed function, restated in oracle/glsl_np.Texture
    def b_texture(self, sampler, uv, bias=None):
        a, _, _ = self.fl(uv)
        return V("vec4", sampler.a.sample(a).astype(F32))

    def b_textureOffset(self, sampler, uv, offset, bias=None):
        a, _, _ = self.fl(uv)
        o = offset.a.astype(np.int64)
        return V("vec4", sampler.a.sample(a, offset=(o[:, 0], o[:, 1])).astype(F32))

    def b_textureGrad(self, sampler, uv, dpdx, dpdy):      # one level: the gradients select nothing
        return self.b_texture(sampler, uv)

    def b_textureProj(self, sampler, p, bias=None):
        a, _, n = self.fl(p)
        return V("vec4", sampler.a.sample((a[:, :2]/a[:, n - 1:n]).astype(F32)).astype(F32))

    def b_texelFetchOffset(self, sampler, p, lod, offset):
        return self.b_texelFetch(sampler, V(p.t, (p.a.astype(np.int64) + offset.a.astype(np.int64)).astype(I32)), lod)

    def b_modf(self, x):
        a, t, _ = self.fl(x)
        whole = np.trunc(a).astype(F32)
        return V(t, (a - whole).astype(F32)), V(t, whole)

    def b_frexp(self, x):
        a, t, n = self.fl(x)
        m, e = np.frexp(a)
        return V(t, m.astype(F32)), V(make_type("int", n), e.astype(I32))

    def b_ldexp(self, x, e):
        a, t, _ = self.fl(x)
        return V(t, np.ldexp(a, e.a.astype(np.int64)).astype(F32))

    def b_textureSize(self, sampler, lod):
        w, h = sampler.a.size
        return V("ivec2", np.asarray([[w, h]], I32))

    def b_texelFetch(self, sampler, p, lod):
        from oracle.glsl_np import texel_fetch
        ix, iy = p.a[:, 0].astype(np.int64), p.a[:, 1].astype(np.int64)
        t = texel_fetch(sampler.a, ix, iy).astype(F32)
        c = t.shape[-1]
        if c < 4:
            pad = [np.zeros(t.shape[:-1] + (1,), F32)]*(3 - c) + [np.ones(t.shape[:-1] + (1,), F32)]
            t = np.concatenate([t] + pad, -1)
        return V("vec4", t)

# ---------------------------------------------------------------------------------------------- #
# The fixed-function pipeline around the two programmable stages

QUAD = ((-1.0, -1.0), (-1.0, 1.0), (1.0, -1.0), (1.0, 1.0))     # shader.py:127-128 (x, y, u, v) = (x, y, x, y)


def interpolate_varyings(vertex_out: dict[str, V], Wr: int, Hr: int, rows=None, cols=None) -> dict[str, np.ndarray]:
    """Fullscreen-quad rasterisation: every `out` of the vertex shader is affine in the vertex position, so the
    value at fragment centre ((i+.5)/Wr, (j+.5)/Hr) is the bilinear blend of the four corner values, computed in
    float64 and rounded once to float32. `flat` integers take the provoking vertex's value. Returns name →
    (n_fragments, dim) arrays in row-major (j, i) order, j = 0 at the bottom."""
    i = (np.arange(Wr, dtype=np.float64) + 0.5)/Wr
    j = (np.arange(Hr, dtype=np.float64) + 0.5)/Hr
    if cols is not None:
        i = i[cols]
    if rows is not None:
        j = j[rows]
    tx, ty = np.meshgrid(i, j)
    tx, ty = tx.reshape(-1), ty.reshape(-1)
    out = {}
    for name, v in vertex_out.items():
        if name.startswith("gl_"):
            continue
        a = np.broadcast_to(v.a, (4,) + v.a.shape[1:])
        if a.dtype != F32:
            out[name] = np.broadcast_to(a[0], (tx.size,) + a.shape[1:]).copy()
            continue
        a = a.astype(np.float64)
        c00, c01, c10, c11 = a[0], a[1], a[2], a[3]        # (x,y) = (-1,-1), (-1,1), (1,-1), (1,1)
        if a.ndim == 1:
            val = (c00*(1 - tx) + c10*tx)*(1 - ty) + (c01*(1 - tx) + c11*tx)*ty
        else:
            val = ((c00[None]*(1 - tx)[:, None] + c10[None]*tx[:, None])*(1 - ty)[:, None]
                   + (c01[None]*(1 - tx)[:, None] + c11[None]*tx[:, None])*ty[:, None])
        out[name] = val.astype(F32)
    return out


class Program:
    """One linked (vertex, fragment) pair as `ShaderProgram.compile` builds it (shader.py:313-349)"""

    def __init__(self, vertex_source: str, fragment_source: str):
        self.vertex = Machine(vertex_source)
        self.fragment = Machine(fragment_source)

    def render(self, uniforms: dict, samplers: dict, Wr: int, Hr: int, rows=None, cols=None) -> np.ndarray:
        """→ fragColor (n_rows, n_cols, 4) float32 before the colour store, row 0 = bottom (or rows[0])"""
        vin = dict(uniforms)
        quad = np.asarray(QUAD, F32)
        vin["vertex_position"] = quad
        vin["vertex_gluv"] = quad
        vs_machine = self.vertex
        vs_machine.outputs.setdefault("gl_Position", "vec4")
        # gl_Position / gl_InstanceID are builtins of the vertex stage
        vs_machine.global_decls = [g for g in vs_machine.global_decls if g[2][0][1] not in ("gl_Position", "gl_InstanceID")]
        vs_machine.global_decls = [("global", [], [("vec4", "gl_Position", None)]),
                                   ("global", [], [("int", "gl_InstanceID", None)])] + vs_machine.global_decls
        vout = vs_machine.run(4, vin, samplers)
        var = interpolate_varyings(vout, Wr, Hr, rows, cols)
        n = next(iter(var.values())).shape[0]
        fin = dict(uniforms)
        fin.update({k: v for k, v in var.items() if k in self.fragment.inputs})
        out = self.fragment.run(n, fin, samplers)
        color = out["fragColor"].a
        color = np.broadcast_to(color, (n, 4))
        nrows = Hr if rows is None else len(np.arange(Hr)[rows])
        return color.reshape(nrows, -1, 4).astype(F32)
