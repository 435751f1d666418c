"""GLSL 3.30 front end of the run-time translator (shaderflow_b200/glsl): tokenizer, preprocessor (#define object- and
function-like with `##`, #undef, #if / #ifdef / #ifndef / #elif / #else / #endif) and a recursive-descent parser to a
tuple AST. It accepts what `ShaderProgram.compile` of the reference hands to the GL driver (shaderflow/shader.py:313-349):
declarations (const / uniform / in / out / flat), structs, overloaded functions with in / out / inout parameters,
arrays, if / for / while / do / switch / break / continue / return / discard and the full expression grammar.

AST nodes (tuples):
    ("function", rtype, name, [(direction, type, name)], ("block", [...]))     ("global", quals, [(type, name, init)])
    statements  ("block", [...]) ("decl", quals, decls) ("expr", e) ("if", c, a, b) ("for", init, cond, step, body)
                ("dowhile", body, c) ("switch", sel, items) ("return", e) ("break",) ("continue",) ("discard",) ("nop",)
    expressions ("lit", type, value) ("name", id) ("call", name, args) ("construct", type, args) ("field", e, name)
                ("index", e, i) ("length", e) ("unary", op, e) ("preinc", op, e) ("postinc", op, e)
                ("binary", op, a, b) ("ternary", c, a, b) ("assign", op, target, value) ("comma", a, b)
    types       a name, or ("array", type, size-expression | None)
"""
from __future__ import annotations

import re
from dataclasses import dataclass

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


SCALARS = ("float", "int", "uint", "bool")
VECTORS = tuple(f"{p}{n}" for p in ("vec", "ivec", "uvec", "bvec") for n in (2, 3, 4))
MATRICES = ("mat2", "mat3", "mat4") + tuple(f"mat{c}x{r}" for c in (2, 3, 4) for r in (2, 3, 4))
BASIC_TYPES = set(SCALARS) | set(VECTORS) | set(MATRICES) | {"void", "sampler2D"}
QUALIFIERS = {"const", "in", "out", "inout", "uniform", "flat", "smooth", "noperspective", "highp", "mediump",
              "lowp", "precise", "centroid", "varying", "attribute"}


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
        if self.peek()[1] == "invariant":              # `invariant gl_Position;` / `invariant out vec4 c;`: no meaning here
            self.next()
            if self.peek(1)[1] == ";":
                self.next(); self.next()
                return ("nop",)
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



def parse(text: str, defines: dict[str, str] | None = None, types: tuple = ()):
    """→ (external declarations, {struct name: [(type, field)]}); `types`: struct names defined outside the text"""
    p = Parser(preprocess(text, defines))
    p.types |= set(types)
    return p.translation_unit(), p.structs
