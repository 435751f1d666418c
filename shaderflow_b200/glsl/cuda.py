"""GLSL AST → CUDA C++ (the back end of the run-time translator).

The emitted translation unit is one `struct Shader : g::ShaderBase` inside namespace `g` (csrc/jit/glsl_rt.cuh,
shaderflow_rt.cuh): GLSL globals become per-thread members, GLSL functions member functions (so they see uniforms,
varyings and each other regardless of order, and a user definition hides a std-lib entry of the same name), `out` /
`inout` parameters references. Types are not inferred here: the C++ templates of glsl_rt.cuh resolve overloads,
int → float promotions and constructor flattening the way GLSL does. What the emitter decides by itself:

    swizzles      `.xyz` / `.rg` … on anything that is not a field of a user struct → swz<…>() / swz_set<…>()
    literals      float literals carry the exact float32 value (GLSL has no double arithmetic)
    members       uniforms and globals are initialised in declaration order when the per-fragment `Shader` is constructed,
                  after ShaderBase has loaded the built-in uniforms and the varyings
    uniforms      names ShaderBase already holds (sfb_uniforms' fixed fields) are dropped; the others get a slot of
                  `extra[16]` or of the sampler table, in order of first declaration and only when the code reads them
    discard       marks the invocation and returns from the current function

`translate()` returns the CUDA source plus the slot tables `ShaderProgram` packs uniforms and samplers by."""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

from shaderflow_b200.glsl import frontend as F

# members of g::ShaderBase (csrc/jit/shaderflow_rt.cuh)
BASE_UNIFORMS = {
    "iTime", "iTau", "iDuration", "iDeltatime", "iResolution", "iWantAspect", "iQuality", "iSSAA", "iFramerate", "iFrame",
    "iRealtime", "iLayer", "iMouseInside", "iMouse", "iMouse1", "iMouse2", "iCameraMode", "iCameraProjection",
    "iCameraPosition", "iCameraRight", "iCameraUpward", "iCameraForward", "iCameraZenith", "iCameraZoom",
    "iCameraIsometric", "iCameraFocalLength", "iCameraOrbital", "iCameraDolly", "iCameraSeparation",
}
BASE_VARYINGS = {"fragCoord", "stxy", "glxy", "stuv", "astuv", "gluv", "agluv", "instance"}
EXTRA_TYPES = set(F.SCALARS) | set(F.VECTORS)
MAX_EXTRA, MAX_SAMPLERS = 16, 20
EXTERNAL_TYPES = ("Camera",)          # structs shaderflow_rt.cuh defines

SWIZZLE_SETS = ("xyzw", "rgba", "stpq")
CXX_RESERVED = {
    "and", "or", "not", "xor", "bitand", "bitor", "compl", "not_eq", "and_eq", "or_eq", "xor_eq", "this", "new", "delete",
    "operator", "template", "typename", "class", "namespace", "using", "typedef", "enum", "union", "auto", "register",
    "extern", "static", "inline", "virtual", "explicit", "friend", "private", "public", "protected", "mutable", "volatile",
    "char", "short", "long", "signed", "unsigned", "double", "goto", "asm", "try", "catch", "throw", "sizeof", "nullptr",
    "alignas", "alignof", "decltype", "constexpr", "noexcept", "static_assert", "thread_local", "wchar_t", "export",
    "near", "far", "min", "max",
} - {"min", "max", "near", "far"}
UNSUPPORTED_CALLS = {"noise1", "noise2", "noise3", "noise4"}      # deprecated since GLSL 1.30; drivers return 0


class TranslationError(RuntimeError):
    pass


@dataclass
class Translation:
    source: str                     # CUDA C++ of `struct Shader` (goes between shaderflow_rt.cuh and jit_kernels.cuh)
    extra: list[str] = field(default_factory=list)        # uniform names → sfb_uniforms.extra[i]
    extra_types: list[str] = field(default_factory=list)
    samplers: list[str] = field(default_factory=list)     # sampler2D names → samplers[i]


def float_literal(value: float) -> str:
    f32 = struct.unpack("f", struct.pack("f", value))[0] if abs(value) < 3.5e38 else value
    if f32 != f32:
        return "__int_as_float(0x7fc00000)"
    if f32 in (float("inf"), float("-inf")) or abs(value) >= 3.5e38:
        return "__int_as_float(0x7f800000)" if value > 0 else "__int_as_float(0xff800000)"
    text = "%.9g" % f32
    if not any(c in text for c in ".en"):
        text += ".0"
    return text + "f"


def ident(name: str) -> str:
    if name == "not":
        return "not_"
    return name + "_" if name in CXX_RESERVED else name


def is_swizzle(name: str) -> bool:
    return 1 <= len(name) <= 4 and any(all(c in s for c in name) for s in SWIZZLE_SETS)


def swizzle_indices(name: str) -> str:
    s = next(s for s in SWIZZLE_SETS if all(c in s for c in name))
    return ", ".join(str(s.index(c)) for c in name)


def is_constant(e, known: set) -> bool:
    """A scalar constant expression C++ can evaluate at compile time (array bounds, case labels need one)"""
    if e is None:
        return False
    kind = e[0]
    if kind == "lit":
        return True
    if kind == "name":
        return e[1] in known
    if kind == "unary":
        return is_constant(e[2], known)
    if kind == "binary":
        return is_constant(e[2], known) and is_constant(e[3], known)
    if kind == "ternary":
        return all(is_constant(x, known) for x in e[1:4])
    if kind == "construct":
        return e[1] in F.SCALARS and len(e[2]) == 1 and is_constant(e[2][0], known)
    return False


def names_used(node, out: set) -> None:
    if isinstance(node, tuple):
        if node and node[0] == "name":
            out.add(node[1])
        elif node and node[0] == "call":
            out.add(node[1])
        for child in node:
            names_used(child, out)
    elif isinstance(node, list):
        for child in node:
            names_used(child, out)


class Emitter:
    def __init__(self, items, structs):
        self.items, self.structs = items, structs
        self.struct_fields = {fname for fields in structs.values() for _, fname in fields}
        self.lines: list[str] = []
        self.depth = 1
        self.rtype = "void"
        # positions of out / inout parameters per user function name (any overload of that arity)
        self.writes: dict[tuple[str, int], set[int]] = {}
        for item in items:
            if item[0] == "function":
                _, _, name, params, _ = item
                slots = self.writes.setdefault((name, len(params)), set())
                slots.update(k for k, (direction, _, _) in enumerate(params) if direction != "in")
        self.function_returns = {(item[2], len(item[3])): item[1] for item in items if item[0] == "function"}
        for builtin in ("modf", "frexp"):                  # the two builtins with an out parameter (GLSL 3.30 / 4.00 §8.3)
            if (builtin, 2) not in self.function_returns:
                self.writes[(builtin, 2)] = {1}
                self.function_returns[(builtin, 2)] = "genType"

    # -- types ----------------------------------------------------------------------------------
    def ctype(self, t, count: int | None = None) -> str:
        """GLSL type → C++ type. Arrays are VALUES in GLSL (assigned, returned, passed by copy), so they become
        g::arr<T, N> (a struct around T[N]) rather than C arrays"""
        if isinstance(t, tuple):
            _, base, size = t
            if isinstance(base, tuple):
                raise TranslationError("arrays of arrays are not supported")
            if size is None and count is None:
                raise TranslationError("unsized array without an initialiser")
            return f"arr<{ident(base)}, {self.expr(size) if size is not None else count}>"
        return ident(t)

    def array_value(self, t, elements) -> str:
        """An array constructor T[n](e0, e1, …) as a braced g::arr; every element converted like a GLSL constructor argument"""
        _, base, size = t
        return self.ctype(t, len(elements)) + "{" + ", ".join(f"{ident(base)}({self.expr(e)})" for e in elements) + "}"

    def declarator(self, t, name: str, init=None) -> tuple[str, str]:
        """→ (C++ declarator 'T name', initialiser text or '')"""
        if not isinstance(t, tuple):
            return f"{self.ctype(t)} {ident(name)}", ("" if init is None else f" = {self.expr(init)}")
        if isinstance(t[1], tuple):
            raise TranslationError(f"'{name}': arrays of arrays are not supported")
        if init is None:
            if t[2] is None:
                raise TranslationError(f"'{name}': unsized array without initialiser")
            return f"{self.ctype(t)} {ident(name)}", ""
        if init[0] == "construct" and isinstance(init[1], tuple):
            count = len(init[2])
            return f"{self.ctype(t, count)} {ident(name)}", " = " + self.array_value((t[0], t[1], t[2] if t[2] is not None else init[1][2]), init[2])
        if t[2] is None:
            raise TranslationError(f"'{name}': an unsized array needs an array constructor as its initialiser")
        return f"{self.ctype(t)} {ident(name)}", f" = {self.expr(init)}"          # any array-valued expression (a call, another array)

    # -- expressions ----------------------------------------------------------------------------
    def expr(self, e) -> str:
        kind = e[0]
        if kind == "lit":
            _, t, v = e
            if t == "float":
                return float_literal(v)
            if t == "bool":
                return "true" if v else "false"
            return f"{v}u" if t == "uint" else str(v)
        if kind == "name":
            return ident(e[1])
        if kind == "call":
            _, name, args = e
            if name in UNSUPPORTED_CALLS:
                raise TranslationError(f"builtin '{name}' is not supported by the CUDA backend")
            # a swizzle handed to an out / inout parameter (`rotate(p.xz, a)`): copy in, call, copy back
            swizzled = [k for k in self.writes.get((name, len(args)), ()) if args[k][0] == "field"
                        and args[k][2] not in self.struct_fields and is_swizzle(args[k][2]) and len(args[k][2]) > 1]
            if swizzled:
                texts = [self.expr(a) for a in args]
                before, after = [], []
                for k in swizzled:
                    base, idx = self.expr(args[k][1]), swizzle_indices(args[k][2])
                    before.append(f"auto sfb_arg{k} = swz<{idx}>({base});")
                    after.append(f"swz_set<{idx}>({base}, sfb_arg{k});")
                    texts[k] = f"sfb_arg{k}"
                call = f"{ident(name)}({', '.join(texts)})"
                if self.function_returns.get((name, len(args)), "void") == "void":
                    return "[&]() { " + " ".join(before) + f" {call}; " + " ".join(after) + " }()"
                return "[&]() { " + " ".join(before) + f" auto sfb_result = {call}; " + " ".join(after) + " return sfb_result; }()"
            return f"{ident(name)}({', '.join(self.expr(a) for a in args)})"
        if kind == "construct":
            _, t, args = e
            if isinstance(t, tuple):
                return self.array_value(t, args)
            return f"{self.ctype(t)}({', '.join(self.expr(a) for a in args)})"
        if kind == "field":
            _, base, name = e
            b = self.expr(base)
            if name not in self.struct_fields and is_swizzle(name) and len(name) > 1:
                return f"swz<{swizzle_indices(name)}>({b})"
            return f"{b}.{ident(name)}"
        if kind == "index":
            return f"{self.expr(e[1])}[{self.expr(e[2])}]"
        if kind == "length":
            return f"length_of({self.expr(e[1])})"
        if kind == "unary":
            return f"({e[1]}{self.expr(e[2])})"
        if kind == "preinc":
            return f"({e[1]}{self.expr(e[2])})"
        if kind == "postinc":
            return f"({self.expr(e[2])}{e[1]})"
        if kind == "binary":
            _, op, a, b = e
            if op == "^^":
                return f"(bool({self.expr(a)}) != bool({self.expr(b)}))"
            return f"({self.expr(a)} {op} {self.expr(b)})"
        if kind == "ternary":
            return f"({self.expr(e[1])} ? {self.expr(e[2])} : {self.expr(e[3])})"
        if kind == "comma":
            return f"({self.expr(e[1])}, {self.expr(e[2])})"
        if kind == "assign":
            _, op, target, value = e
            v = self.expr(value)
            if target[0] == "field" and target[2] not in self.struct_fields and is_swizzle(target[2]) and len(target[2]) > 1:
                base, idx = self.expr(target[1]), swizzle_indices(target[2])
                if op == "=":
                    return f"swz_set<{idx}>({base}, {v})"
                return f"swz_set<{idx}>({base}, swz<{idx}>({base}) {op[:-1]} ({v}))"
            return f"({self.expr(target)} {op} {v})"
        raise TranslationError(f"unsupported expression node {kind!r}")

    # -- statements -----------------------------------------------------------------------------
    def put(self, text: str) -> None:
        self.lines.append("    "*self.depth + text)

    def decl(self, quals, decls) -> None:
        prefix = "const " if "const" in quals else ""
        for (t, name, init) in decls:
            d, i = self.declarator(t, name, init)
            if init is not None:
                mentioned: set = set()
                names_used(init, mentioned)
                if name in mentioned:
                    # `vec2 gluv = gluv - offset;` (camera.glsl:101): in GLSL a name is not in scope in its own initialiser,
                    # so this reads the OUTER gluv; in C++ it would read the new, uninitialised one
                    self.shadows = getattr(self, "shadows", 0) + 1
                    held = f"sfb_outer{self.shadows}"
                    self.put(f"const auto {held}{i};")
                    i = f" = {held}"
            self.put(f"{prefix}{d}{i};")

    def block(self, node) -> None:
        """a statement as the body of a control structure: always braced"""
        self.put("{")
        self.depth += 1
        for s in (node[1] if node[0] == "block" else [node]):
            self.stmt(s)
        self.depth -= 1
        self.put("}")

    def stmt(self, s) -> None:
        kind = s[0]
        if kind == "nop":
            return
        if kind == "block":
            self.block(s)
        elif kind == "decl":
            self.decl(s[1], s[2])
        elif kind == "expr":
            self.put(self.expr(s[1]) + ";")
        elif kind == "if":
            self.put(f"if ({self.expr(s[1])})")
            self.block(s[2])
            if s[3] is not None:
                self.put("else")
                self.block(s[3])
        elif kind == "for":
            _, init, cond, step, body = s
            self.put("{")
            self.depth += 1
            self.stmt(init)
            self.put(f"for (; {self.expr(cond) if cond is not None else ''}; {self.expr(step) if step is not None else ''})")
            self.block(body)
            self.depth -= 1
            self.put("}")
        elif kind == "dowhile":
            self.put("do")
            self.block(s[1])
            self.put(f"while ({self.expr(s[2])});")
        elif kind == "switch":
            self.put(f"switch ({self.expr(s[1])}) {{")
            self.depth += 1
            for item in s[2]:
                if item[0] == "case":
                    self.put(f"case {self.expr(item[1])}:")
                elif item[0] == "default":
                    self.put("default:")
                else:
                    self.stmt(item)
            self.depth -= 1
            self.put("}")
        elif kind == "return":
            self.put("return;" if s[1] is None else f"return {self.expr(s[1])};")
        elif kind in ("break", "continue"):
            self.put(kind + ";")
        elif kind == "discard":
            self.put("{ sfb_discarded = true; return" + ("" if self.rtype == "void" else f" {self.ctype(self.rtype)}()") + "; }")
        else:
            raise TranslationError(f"unsupported statement node {kind!r}")

    # -- top level ------------------------------------------------------------------------------
    def translate(self) -> Translation:
        # what main() reaches: the assembled header declares every module's uniforms, texture accessors and the whole
        # std-lib, and only what is read may take a uniform / sampler slot
        bodies: dict[str, list] = {}
        for item in self.items:
            if item[0] == "function":
                bodies.setdefault(item[2], []).append((item[3], item[4]))
            elif item[0] == "global":
                for (t, name, init) in item[2]:
                    bodies.setdefault(name, []).append((t, init))
        used: set = set()
        frontier = ["main"]
        while frontier:
            name = frontier.pop()
            if name in used:
                continue
            used.add(name)
            found: set = set()
            names_used(bodies.get(name, []), found)
            frontier.extend(found - used)
        out = Translation(source="")
        functions = {item[2] for item in self.items if item[0] == "function"}
        if "main" not in functions:
            raise TranslationError("the fragment shader has no main()")

        self.put("// structs")
        for sname, fields in self.structs.items():
            self.put(f"struct {ident(sname)} {{")
            self.depth += 1
            plain = []
            for (ft, fname) in fields:
                d, _ = self.declarator(ft, fname)
                self.put(d + ";")
                plain.append((self.ctype(ft), ident(fname)))
            if len(plain) == len(fields) and fields:
                self.put(f"G_DEV {ident(sname)}() {{}}")
                params = ", ".join(f"{t} {n}_" for t, n in plain)
                inits = ", ".join(f"{n}({n}_)" for _, n in plain)
                self.put(f"G_DEV {ident(sname)}({params}) : {inits} {{}}")
            self.depth -= 1
            self.put("};")

        self.put("// uniforms, globals")
        declared: set = set()
        constants: set = set()
        for item in self.items:
            if item[0] != "global":
                continue
            _, quals, decls = item
            for (t, name, init) in decls:
                if name in declared:
                    continue
                if "uniform" in quals:
                    if name in BASE_UNIFORMS:
                        continue
                    if name not in used:
                        continue                               # the header declares every module's uniforms
                    declared.add(name)
                    if t == "sampler2D":
                        if len(out.samplers) >= MAX_SAMPLERS:
                            raise TranslationError(f"more than {MAX_SAMPLERS} samplers are read by the shader")
                        self.put(f"sampler2D {ident(name)} = sfb_sampler({len(out.samplers)});")
                        out.samplers.append(name)
                    elif t in EXTRA_TYPES:
                        if len(out.extra) >= MAX_EXTRA:
                            raise TranslationError(f"more than {MAX_EXTRA} module / user uniforms are read by the shader")
                        self.put(f"{self.ctype(t)} {ident(name)} = sfb_extra<{self.ctype(t)}>({len(out.extra)});")
                        out.extra.append(name)
                        out.extra_types.append(t)
                    elif t in ("mat2", "mat3", "mat4"):
                        # one slot per column (variable.py's GlslType lists the three square matrices); the value
                        # arrives as GL takes it: n*n floats, column after column
                        columns = int(t[3])
                        if len(out.extra) + columns > MAX_EXTRA:
                            raise TranslationError(f"more than {MAX_EXTRA} slots of module / user uniforms are read by the shader")
                        self.put(f"{t} {ident(name)} = sfb_extra_matrix<{columns}>({len(out.extra)});")
                        for column in range(columns):
                            out.extra.append(name)
                            out.extra_types.append(f"{t}:{column}")
                    else:
                        raise TranslationError(f"uniform '{name}' of type {t} is not supported (scalars, vectors, square matrices and sampler2D are)")
                elif "in" in quals or "varying" in quals:
                    if name in BASE_VARYINGS:
                        continue
                    if name in used:
                        raise TranslationError(f"varying '{name}' is not produced by the fullscreen vertex stage")
                elif "out" in quals:
                    declared.add(name)
                    if name != "fragColor":
                        self.put(f"{self.ctype(t)}& {ident(name)} = fragColor;")      # the one colour attachment
                else:
                    if name not in used:
                        continue
                    declared.add(name)
                    d, i = self.declarator(t, name, init)
                    if "const" in quals and t in F.SCALARS and is_constant(init, constants):
                        constants.add(name)                    # usable as an array bound or a case label
                        self.put(f"static constexpr {d}{i};")
                    else:
                        self.put(("const " if "const" in quals else "") + d + i + ";")

        self.put("// functions")
        for item in self.items:
            if item[0] != "function":
                continue
            _, rtype, name, params, body = item
            if name not in used:
                continue
            self.rtype = rtype
            plist = []
            for (direction, pt, pname) in params:
                pname = pname or f"sfb_unnamed{len(plist)}"
                plist.append(f"{self.ctype(pt)}{'&' if direction != 'in' else ''} {ident(pname)}")
            self.put(f"G_DEV {self.ctype(rtype)} {ident(name)}({', '.join(plist)})")
            self.block(body)
        self.put("G_DEV Shader(const RenderParams& P, int i, int j) : ShaderBase(P, i, j) {}")
        out.source = "namespace g {\nstruct Shader : ShaderBase {\n" + "\n".join(self.lines) + "\n};\n}  // namespace g\n"
        return out


# GLSL the translator puts in front of every fragment: the function-like macros of the std-lib API
PRELUDE = """
#define GetCamera(name) Camera name = sfb_get_camera()
"""


def translate(fragment: str, header: str = "", defines: dict[str, str] | None = None) -> Translation:
    """fragment: the user's GLSL (`main()` and whatever it declares); header: declarations in front of it (uniforms of
    the pipeline, texture aliases — what shader.py:190-239 assembles). → Translation"""
    try:
        items, structs = F.parse(PRELUDE + header + "\n" + fragment, defines, types=EXTERNAL_TYPES)
    except SyntaxError as error:
        raise TranslationError(str(error)) from None
    return Emitter(items, structs).translate()
