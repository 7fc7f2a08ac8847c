"""tsinterp — a small interpreter for the TypeScript subset the reference's `src/*.ts` is written in.

WHY.  The oracle (oracle/bbq_oracle.cpp) is a hand restatement of the reference's numerics.  Nothing in this image or on
the GPU box can run JavaScript / TypeScript (profiles/r02_js_runtime_probe.txt), so the restatement could only be pinned
by reading.  This module executes the reference's UNMODIFIED source files (read from /root/reference at generation
time, never copied) with JavaScript semantics — IEEE doubles, Float32Array / Uint8Array / Int32Array stores, ToInt32
bit operations, `Math.round` half-up, NaN-propagating `Math.min/max`, strict left-to-right evaluation — so that golden
vectors can be produced FROM THE REFERENCE'S OWN TEXT (make_golden_with_interp.py) and compared with the oracle
(tests/test_golden_from_ts.py).  It is test infrastructure: nothing in the product imports it.

WHAT.  A lexer, a recursive-descent parser that understands (and discards) TypeScript's type syntax — annotations,
generics, interfaces, type aliases, `as`, non-null `!`, access modifiers, parameter properties, enums — and a compiler
from the AST to Python closures.  Supported: ES modules (named imports / exports, re-exports), classes (fields, static
members, methods, constructors), functions / arrow functions / closures, destructuring, spread, template literals,
`for` / `for-of` / `for-in` / `while` / `do` / `switch` / `try`, the operators of ES2020, typed arrays, `Array`, `Map`,
`Set`, `Math`, `Number`, `Object`, `JSON.stringify` (plain data), `Error`.  Not supported (unused by the reference):
generators, async, getters / setters, inheritance, labels, regular expressions, tagged templates, `with`, `eval`.
"""
from __future__ import annotations

import math
import os
import re
import struct
import sys
from array import array


# ---------------------------------------------------------------------------------------------------------------
# values
# ---------------------------------------------------------------------------------------------------------------
class Undefined:
    __slots__ = ()

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


UNDEF = Undefined()


class JSThrow(Exception):
    def __init__(self, value):
        Exception.__init__(self)
        self.value = value

    def __str__(self):
        v = self.value
        if isinstance(v, JSObj) and "message" in v.props:
            return f"{v.props.get('name', 'Error')}: {v.props['message']}"
        return to_str(v)


class JSObj:
    __slots__ = ("cls", "props")

    def __init__(self, cls, props=None):
        self.cls = cls
        self.props = {} if props is None else props


class JSClass:
    __slots__ = ("name", "ctor", "methods", "statics", "fields", "native_new", "parent")

    def __init__(self, name):
        self.name, self.ctor, self.methods, self.statics, self.fields, self.native_new = name, None, {}, {}, [], None
        self.parent = None

    def find_method(self, key):
        c = self
        while c is not None:
            f = c.methods.get(key)
            if f is not None:
                return f
            c = c.parent
        return None

    def find_static(self, key):
        c = self
        while c is not None:
            if key in c.statics:
                return c.statics[key]
            c = c.parent
        return UNDEF


class JSRegExp:
    __slots__ = ("source", "flags", "rx")

    def __init__(self, source, flags):
        f = 0
        for ch in flags:
            f |= {"i": re.I, "m": re.M, "s": re.S}.get(ch, 0)
        self.source, self.flags, self.rx = source, flags, re.compile(source.replace("(?<", "(?P<"), f)


class JSFunction:
    __slots__ = ("params", "body", "env", "is_arrow", "is_expr", "name", "simple", "param_props")

    def __init__(self, params, body, env, is_arrow, is_expr, name, simple, param_props):
        self.params, self.body, self.env = params, body, env
        self.is_arrow, self.is_expr, self.name, self.simple, self.param_props = is_arrow, is_expr, name, simple, param_props


class BoundMethod:  # obj.method read as a value (not called on the spot)
    __slots__ = ("this", "fn")

    def __init__(self, this, fn):
        self.this, self.fn = this, fn


KIND_CODE = {"Float32Array": "f", "Float64Array": "d", "Uint8Array": "B", "Int8Array": "b", "Int32Array": "i",
             "Uint32Array": "I", "Uint16Array": "H", "Int16Array": "h", "Uint8ClampedArray": "B"}
INT_BITS = {"B": (8, False), "b": (8, True), "i": (32, True), "I": (32, False), "H": (16, False), "h": (16, True)}


class TypedArray:
    __slots__ = ("kind", "a", "clamped")

    def __init__(self, kind, a):
        self.kind, self.a, self.clamped = kind, a, kind == "Uint8ClampedArray"

    def store(self, i, v):
        code = self.a.typecode
        x = v if type(v) is float else to_number(v)
        if code == "f" or code == "d":
            self.a[i] = x
            return
        if x != x or x in (math.inf, -math.inf):
            self.a[i] = 0
            return
        bits, signed = INT_BITS[code]
        if self.clamped:
            self.a[i] = 0 if x < 0 else 255 if x > 255 else int(round_half_even(x))
            return
        n = int(x) & ((1 << bits) - 1)
        if signed and n >= (1 << (bits - 1)):
            n -= 1 << bits
        self.a[i] = n


def round_half_even(x):
    return round(x)


class JSMap:
    __slots__ = ("d",)

    def __init__(self):
        self.d = {}


class JSSet:
    __slots__ = ("d",)

    def __init__(self):
        self.d = {}


class Native:  # a host function / namespace: callable and / or with properties, optionally constructible
    __slots__ = ("name", "call", "props", "construct")

    def __init__(self, name, call=None, props=None, construct=None):
        self.name, self.call, self.props, self.construct = name, call, props or {}, construct


def map_key(v):
    if type(v) is float:
        if v != v:
            return ("nan",)
        return ("n", v + 0.0)
    if isinstance(v, (str, bool)) or v is None or v is UNDEF:
        return (type(v).__name__, v)
    return ("o", id(v))


# ---------------------------------------------------------------------------------------------------------------
# conversions and operators (ECMAScript semantics)
# ---------------------------------------------------------------------------------------------------------------
def truthy(v):
    t = type(v)
    if t is bool:
        return v
    if t is float:
        return v == v and v != 0.0
    if v is None or v is UNDEF:
        return False
    if t is str:
        return len(v) > 0
    return True


def to_number(v):
    t = type(v)
    if t is float:
        return v
    if t is bool:
        return 1.0 if v else 0.0
    if v is None:
        return 0.0
    if v is UNDEF:
        return math.nan
    if t is str:
        s = v.strip()
        if s == "":
            return 0.0
        try:
            if s[:2].lower() == "0x":
                return float(int(s, 16))
            if s in ("Infinity", "+Infinity"):
                return math.inf
            if s == "-Infinity":
                return -math.inf
            return float(s) if re.fullmatch(r"[+-]?(\d+\.?\d*(e[+-]?\d+)?|\.\d+(e[+-]?\d+)?)", s, re.I) else math.nan
        except ValueError:
            return math.nan
    if t is int:
        return float(v)
    if isinstance(v, list):
        return to_number(to_str(v))
    return math.nan


def to_int32(v):
    x = v if type(v) is float else to_number(v)
    if x != x or x == math.inf or x == -math.inf:
        return 0
    n = int(x) & 0xFFFFFFFF
    return n - 0x100000000 if n >= 0x80000000 else n


def to_uint32(v):
    x = v if type(v) is float else to_number(v)
    if x != x or x == math.inf or x == -math.inf:
        return 0
    return int(x) & 0xFFFFFFFF


def num_to_str(x):
    if x != x:
        return "NaN"
    if x == math.inf:
        return "Infinity"
    if x == -math.inf:
        return "-Infinity"
    if x == int(x) and abs(x) < 1e21:
        return str(int(x))
    r = repr(x)
    if "e" in r:
        m, e = r.split("e")
        sign = "-" if e[0] == "-" else "+"
        r = f"{m}e{sign}{int(e.lstrip('+-'))}"
    return r


def to_str(v):
    t = type(v)
    if t is str:
        return v
    if t is float:
        return num_to_str(v)
    if t is bool:
        return "true" if v else "false"
    if v is None:
        return "null"
    if v is UNDEF:
        return "undefined"
    if t is list:
        return ",".join("" if (e is None or e is UNDEF) else to_str(e) for e in v)
    if t is TypedArray:
        return ",".join(num_to_str(float(e)) for e in v.a)
    if t is JSObj:
        if "message" in v.props and v.cls is not None and v.cls.name.endswith("Error"):
            return f"{v.props.get('name', v.cls.name)}: {to_str(v.props['message'])}"
        return "[object Object]"
    if t is dict:
        return "[object Object]"
    if t in (JSFunction, Native, BoundMethod, JSClass):
        return "function"
    return str(v)


def prop_key(v):
    if type(v) is str:
        return v
    return to_str(v)


def js_typeof(v):
    t = type(v)
    if t is float:
        return "number"
    if t is str:
        return "string"
    if t is bool:
        return "boolean"
    if v is UNDEF:
        return "undefined"
    if t in (JSFunction, BoundMethod, JSClass) or (t is Native and (v.call or v.construct)):
        return "function"
    return "object"


def strict_eq(a, b):
    ta, tb = type(a), type(b)
    if ta is float and tb is float:
        return a == b
    if ta is not tb:
        return False
    if ta in (str, bool):
        return a == b
    return a is b


def loose_eq(a, b):
    if (a is None or a is UNDEF) and (b is None or b is UNDEF):
        return True
    if a is None or a is UNDEF or b is None or b is UNDEF:
        return False
    ta, tb = type(a), type(b)
    if ta is tb or (ta is float and tb is float):
        return strict_eq(a, b)
    if ta in (float, str, bool) and tb in (float, str, bool):
        return to_number(a) == to_number(b)
    return a is b


def js_add(a, b):
    if type(a) is float and type(b) is float:
        return a + b
    if isinstance(a, (list, dict, JSObj, TypedArray)):
        a = to_str(a)
    if isinstance(b, (list, dict, JSObj, TypedArray)):
        b = to_str(b)
    if type(a) is str or type(b) is str:
        return to_str(a) + to_str(b)
    return to_number(a) + to_number(b)


def js_div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


def js_mod(a, b):
    try:
        return math.fmod(a, b)
    except (ValueError, ZeroDivisionError):
        return math.nan


def js_pow(a, b):
    if b != b:
        return math.nan
    if b == 0.0:
        return 1.0
    if (a == 1.0 or a == -1.0) and (b == math.inf or b == -math.inf):
        return math.nan
    try:
        r = math.pow(a, b)
    except OverflowError:
        return math.inf if (a > 0 or b % 2 == 0) else -math.inf
    except (ValueError, ZeroDivisionError):
        if a == 0.0:
            return math.inf if (b % 2 == 0 or math.copysign(1.0, a) > 0) else -math.inf
        return math.nan
    return r


def js_compare(op, a, b):
    if type(a) is str and type(b) is str:
        pass
    else:
        a, b = to_number(a), to_number(b)
    if op == "<":
        return a < b
    if op == ">":
        return a > b
    if op == "<=":
        return a <= b
    return a >= b


def math_round(x):
    x = to_number(x)
    if x != x or x == math.inf or x == -math.inf:
        return x
    f = math.floor(x)
    r = f + 1.0 if x - f >= 0.5 else f
    if r == 0.0 and (x < 0.0 or math.copysign(1.0, x) < 0):
        return -0.0
    return float(r)


def math_max(*args):
    r = -math.inf
    nan = False
    for v in args:
        x = to_number(v)
        if x != x:
            nan = True
        elif x > r or (x == 0.0 and r == 0.0 and math.copysign(1.0, x) > 0):
            r = x
    return math.nan if nan else r


def math_min(*args):
    r = math.inf
    nan = False
    for v in args:
        x = to_number(v)
        if x != x:
            nan = True
        elif x < r or (x == 0.0 and r == 0.0 and math.copysign(1.0, x) < 0):
            r = x
    return math.nan if nan else r


def _m1(fn):
    def f(x=UNDEF):
        x = to_number(x)
        try:
            return float(fn(x))
        except (ValueError, OverflowError):
            if fn in (math.exp, math.cosh, math.sinh, math.expm1) and x == x:
                return math.inf if (x > 0 or fn is math.cosh) else (-math.inf if fn is math.sinh else 0.0)
            return math.nan
    return f


def _floor(x):
    return x if (x != x or x in (math.inf, -math.inf)) else (float(math.floor(x)) if x != 0 else x)


def _ceil(x):
    if x != x or x in (math.inf, -math.inf) or x == 0:
        return x
    r = float(math.ceil(x))
    return -0.0 if (r == 0.0 and x < 0) else r


def _trunc(x):
    if x != x or x in (math.inf, -math.inf) or x == 0:
        return x
    r = float(math.trunc(x))
    return -0.0 if (r == 0.0 and x < 0) else r


def _sqrt(x):
    return math.sqrt(x) if x >= 0 else (x if x == 0 else math.nan)


def _log(fn):
    def f(x=UNDEF):
        x = to_number(x)
        if x != x or x < 0:
            return math.nan
        if x == 0:
            return -math.inf
        if x == math.inf:
            return math.inf
        return float(fn(x))
    return f


def _fround(x=UNDEF):
    x = to_number(x)
    if x != x or x in (math.inf, -math.inf):
        return x
    try:
        return struct.unpack("f", struct.pack("f", x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def _sign(x=UNDEF):
    x = to_number(x)
    return x if (x != x or x == 0) else (1.0 if x > 0 else -1.0)


def _hypot(*a):
    v = [to_number(x) for x in a]
    if any(x in (math.inf, -math.inf) for x in v):
        return math.inf
    if any(x != x for x in v):
        return math.nan
    return math.hypot(*v) if v else 0.0


# ---------------------------------------------------------------------------------------------------------------
# lexer
# ---------------------------------------------------------------------------------------------------------------
class Tok:
    __slots__ = ("kind", "val", "pos", "nl")

    def __init__(self, kind, val, pos, nl):
        self.kind, self.val, self.pos, self.nl = kind, val, pos, nl

    def __repr__(self):
        return f"{self.kind}:{self.val!r}"


PUNCS = [">>>=", "...", "===", "!==", "**=", "<<=", ">>=", ">>>", "&&=", "||=", "??=", "=>", "==", "!=", "<=", ">=", "&&",
         "||", "??", "?.", "++", "--", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "**", "<<", ">>"]
TOKEN_RE = re.compile(
    r"(?P<ws>[ \t\r\f\v]+)|(?P<nl>\n)|(?P<lc>//[^\n]*)|(?P<bc>/\*.*?\*/)|"
    r"(?P<num>0[xX][0-9a-fA-F_]+|0[bB][01_]+|0[oO][0-7_]+|(?:\d[\d_]*\.?[\d_]*|\.\d[\d_]*)(?:[eE][+-]?\d+)?)|"
    r"(?P<id>[A-Za-z_$\u0080-\uffff][\w$\u0080-\uffff]*)|"
    r"(?P<str>'(?:\\.|[^'\\\n])*'|\"(?:\\.|[^\"\\\n])*\")|"
    r"(?P<punc>" + "|".join(re.escape(p) for p in PUNCS) + r"|[{}()\[\];,<>+\-*/%&|^!~?:=.@#])", re.S)
ESCAPES = {"n": "\n", "t": "\t", "r": "\r", "b": "\b", "f": "\f", "v": "\v", "0": "\0", "\n": ""}


def unescape(s):
    out, i = [], 0
    while i < len(s):
        c = s[i]
        if c != "\\":
            out.append(c)
            i += 1
            continue
        c = s[i + 1]
        if c == "u":
            if s[i + 2] == "{":
                j = s.index("}", i)
                out.append(chr(int(s[i + 3:j], 16)))
                i = j + 1
            else:
                out.append(chr(int(s[i + 2:i + 6], 16)))
                i += 6
        elif c == "x":
            out.append(chr(int(s[i + 2:i + 4], 16)))
            i += 4
        else:
            out.append(ESCAPES.get(c, c))
            i += 2
    return "".join(out)


def lex(src, fname="<ts>"):
    toks, i, n, nl = [], 0, len(src), False
    while i < n:
        if src[i] == "`":  # template literal: raw scan with ${ } nesting
            j, parts, cur = i + 1, [], []
            while True:
                if j >= n:
                    raise SyntaxError(f"{fname}: unterminated template literal at {i}")
                c = src[j]
                if c == "`":
                    break
                if c == "\\":
                    cur.append(src[j:j + 2])
                    j += 2
                elif c == "$" and src[j + 1] == "{":
                    parts.append(unescape("".join(cur)))
                    cur = []
                    depth, k = 1, j + 2
                    while depth:
                        if src[k] == "{":
                            depth += 1
                        elif src[k] == "}":
                            depth -= 1
                        elif src[k] in "'\"`":  # (nested strings containing braces: not used by the reference)
                            q = src[k]
                            k += 1
                            while src[k] != q:
                                k += 2 if src[k] == "\\" else 1
                        k += 1
                    parts.append(lex(src[j + 2:k - 1], fname))
                    j = k
                else:
                    cur.append(c)
                    j += 1
            parts.append(unescape("".join(cur)))
            toks.append(Tok("tpl", parts, i, nl))
            nl = False
            i = j + 1
            continue
        if src[i] == "/" and src[i + 1:i + 2] not in ("/", "*"):   # division or a regular-expression literal?
            prev = toks[-1] if toks else None
            postfix_bang = prev is not None and prev.kind == "punc" and prev.val == "!" and len(toks) >= 2 and \
                (toks[-2].kind in ("id", "num", "str", "tpl") or (toks[-2].kind == "punc" and toks[-2].val in (")", "]")))   # x! / y
            starts = prev is None or (prev.kind == "punc" and prev.val not in (")", "]", "}") and not postfix_bang) or \
                (prev.kind == "id" and prev.val in ("return", "typeof", "case", "do", "else", "in", "of", "void", "delete", "throw", "new"))
            if starts:
                j, in_class = i + 1, False
                while True:
                    if j >= n or src[j] == "\n":
                        raise SyntaxError(f"{fname}: unterminated regular expression at {i}")
                    c = src[j]
                    if c == "\\":
                        j += 2
                        continue
                    if c == "[":
                        in_class = True
                    elif c == "]":
                        in_class = False
                    elif c == "/" and not in_class:
                        break
                    j += 1
                k = j + 1
                while k < n and src[k].isalpha():
                    k += 1
                toks.append(Tok("regex", (src[i + 1:j], src[j + 1:k]), i, nl))
                nl = False
                i = k
                continue
        m = TOKEN_RE.match(src, i)
        if not m:
            raise SyntaxError(f"{fname}: cannot tokenise at {i}: {src[i:i + 30]!r}")
        kind = m.lastgroup
        text = m.group()
        i = m.end()
        if kind == "nl" or (kind == "bc" and "\n" in text):
            nl = True
            continue
        if kind in ("ws", "lc", "bc"):
            continue
        if kind == "num":
            t = text.replace("_", "")
            v = float(int(t, 16)) if t[:2] in ("0x", "0X") else float(int(t[2:], 2)) if t[:2] in ("0b", "0B") else \
                float(int(t[2:], 8)) if t[:2] in ("0o", "0O") else float(t)
            toks.append(Tok("num", v, m.start(), nl))
        elif kind == "str":
            toks.append(Tok("str", unescape(text[1:-1]), m.start(), nl))
        else:
            toks.append(Tok(kind, text, m.start(), nl))
        nl = False
    toks.append(Tok("eof", None, n, True))
    return toks


# ---------------------------------------------------------------------------------------------------------------
# parser (TypeScript syntax in, plain-JavaScript AST out: every type construct is skipped here)
# ---------------------------------------------------------------------------------------------------------------
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "**=", "<<=", ">>=", ">>>=", "&=", "|=", "^=", "&&=", "||=", "??="}
BIN_PREC = {"??": 1, "||": 2, "&&": 3, "|": 4, "^": 5, "&": 6, "==": 7, "!=": 7, "===": 7, "!==": 7, "<": 8, ">": 8, "<=": 8,
            ">=": 8, "instanceof": 8, "in": 8, "<<": 9, ">>": 9, ">>>": 9, "+": 10, "-": 10, "*": 11, "/": 11, "%": 11, "**": 12}
MODIFIERS = {"public", "private", "protected", "readonly", "static", "abstract", "override", "declare"}
RESERVED_STARTS = {"function", "class", "new", "this", "null", "true", "false", "undefined", "typeof", "void", "delete"}


class Parser:
    def __init__(self, toks, fname):
        self.t, self.i, self.fname = toks, 0, fname

    # -- token helpers
    def peek(self, k=0):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def at(self, v, k=0):
        tok = self.peek(k)
        return tok.kind == "punc" and tok.val == v

    def at_id(self, v, k=0):
        tok = self.peek(k)
        return tok.kind == "id" and tok.val == v

    def eat(self, v):
        if self.at(v):
            self.i += 1
            return True
        return False

    def eat_id(self, v):
        if self.at_id(v):
            self.i += 1
            return True
        return False

    def fail(self, msg):
        tok = self.peek()
        raise SyntaxError(f"{self.fname}: {msg} at token {tok!r} (offset {tok.pos})")

    def expect(self, v):
        if not self.eat(v):
            self.fail(f"expected {v!r}")

    def ident(self):
        tok = self.next()
        if tok.kind != "id":
            self.i -= 1
            self.fail("expected identifier")
        return tok.val

    def semi(self):
        if self.eat(";"):
            return
        tok = self.peek()
        if tok.kind == "eof" or self.at("}") or tok.nl:
            return
        self.fail("expected ';'")

    # -- types (skipped)
    def skip_balanced(self, open_, close):
        self.expect(open_)
        depth = 1
        while depth:
            tok = self.next()
            if tok.kind == "eof":
                self.fail(f"unbalanced {open_}")
            if tok.kind == "punc":
                if tok.val == open_:
                    depth += 1
                elif tok.val == close:
                    depth -= 1

    def skip_type_args(self):  # at '<': returns False (position restored) if this is not a well-formed <...>
        start = self.i
        if not self.at("<"):
            return False
        depth = 0
        while True:
            tok = self.next()
            if tok.kind == "eof":
                self.i = start
                return False
            if tok.kind == "punc":
                v = tok.val
                if v == "<":
                    depth += 1
                elif v == ">":
                    depth -= 1
                elif v == ">>":
                    depth -= 2
                elif v == ">>>":
                    depth -= 3
                elif v in ("(", "[", "{"):
                    self.i -= 1
                    self.skip_balanced(v, {"(": ")", "[": "]", "{": "}"}[v])
                elif v in (")", "]", "}", ";", "&&", "||", "+", "-", "*", "/", "==", "===", "!=", "!==", "++", "--", "!", "<=", ">="):
                    self.i = start
                    return False
                if depth <= 0:
                    return depth == 0 or self._unsplit(start)
            elif tok.kind == "num" and depth > 0 and not (self.t[self.i - 2].kind == "punc" and self.t[self.i - 2].val in ("<", ",", "|", "[")):
                self.i = start
                return False

    def _unsplit(self, start):
        self.i = start
        return False

    def skip_type(self):
        self.eat("|")
        self.eat("&")
        while True:
            self.skip_type_primary()
            if self.at("|") or self.at("&"):
                self.i += 1
                continue
            if self.at_id("extends") and False:
                continue
            break

    def skip_type_primary(self):
        tok = self.peek()
        if tok.kind == "id" and tok.val in ("keyof", "typeof", "readonly", "unique", "infer", "asserts", "new", "abstract"):
            self.i += 1
            return self.skip_type_primary()
        if self.at("("):
            self.skip_balanced("(", ")")
            if self.eat("=>"):
                self.skip_type()
                return
        elif self.at("{"):
            self.skip_balanced("{", "}")
        elif self.at("["):
            self.skip_balanced("[", "]")
        elif self.at("<"):  # generic function type  <T>(x: T) => U
            self.skip_type_args()
            return self.skip_type_primary()
        elif tok.kind in ("str", "num", "tpl"):
            self.i += 1
        elif self.at("-") and self.peek(1).kind == "num":
            self.i += 2
        elif tok.kind == "id":
            self.i += 1
            while self.at(".") and self.peek(1).kind == "id":
                self.i += 2
            if self.at("<"):
                if not self.skip_type_args():
                    self.fail("malformed type arguments")
            if self.at_id("is"):
                self.i += 1
                self.skip_type()
                return
        else:
            self.fail("type expected")
        while self.at("[") and not self.peek().nl:  # T[] / T[K]
            self.skip_balanced("[", "]")

    def opt_type_annotation(self):
        if self.eat(":"):
            self.skip_type()

    # -- module items and statements
    def program(self):
        body = []
        while self.peek().kind != "eof":
            body.append(self.statement(top=True))
        return body

    def statement(self, top=False):
        tok = self.peek()
        if tok.kind == "punc":
            if tok.val == "{":
                return ("block", self.block())
            if tok.val == ";":
                self.i += 1
                return ("empty",)
        if tok.kind == "id":
            v = tok.val
            nxt = self.peek(1)
            if v == "import" and not (nxt.kind == "punc" and nxt.val in ("(", ".")):
                return self.import_decl()
            if v == "export":
                return self.export_decl()
            if v in ("const", "let", "var") and not (v == "let" and nxt.kind == "punc" and nxt.val not in ("[", "{")):
                if v == "const" and self.at_id("enum", 1):
                    self.i += 1
                    return self.enum_decl()
                d = self.var_decl()
                self.semi()
                return d
            if v == "function":
                return self.function_decl()
            if v == "class" or (v == "abstract" and self.at_id("class", 1)):
                self.eat_id("abstract")
                return self.class_decl()
            if v == "enum" and nxt.kind == "id":
                return self.enum_decl()
            if v == "interface" and nxt.kind == "id":
                self.i += 2
                while not self.at("{"):
                    self.i += 1
                self.skip_balanced("{", "}")
                return ("empty",)
            if v == "type" and nxt.kind == "id" and (self.at("=", 2) or self.at("<", 2)):
                self.skip_type_alias()
                return ("empty",)
            if v == "declare":
                self.i += 1
                self.statement()
                return ("empty",)
            if v == "if":
                self.i += 1
                self.expect("(")
                c = self.expression()
                self.expect(")")
                a = self.statement()
                b = self.statement() if self.eat_id("else") else None
                return ("if", c, a, b)
            if v == "for":
                return self.for_stmt()
            if v == "while":
                self.i += 1
                self.expect("(")
                c = self.expression()
                self.expect(")")
                return ("while", c, self.statement())
            if v == "do":
                self.i += 1
                body = self.statement()
                if not self.eat_id("while"):
                    self.fail("expected while")
                self.expect("(")
                c = self.expression()
                self.expect(")")
                self.semi()
                return ("dowhile", body, c)
            if v == "return":
                self.i += 1
                e = None
                if not (self.at(";") or self.at("}") or self.peek().nl or self.peek().kind == "eof"):
                    e = self.expression()
                self.semi()
                return ("ret", e)
            if v == "throw":
                self.i += 1
                e = self.expression()
                self.semi()
                return ("throw", e)
            if v == "break" or v == "continue":
                self.i += 1
                if self.peek().kind == "id" and not self.peek().nl:
                    self.fail("labels are not supported")
                self.semi()
                return ("break",) if v == "break" else ("cont",)
            if v == "switch":
                return self.switch_stmt()
            if v == "try":
                return self.try_stmt()
        e = self.expression()
        self.semi()
        return ("expr", e)

    def skip_type_alias(self):
        self.i += 2
        depth = 0
        while True:
            tok = self.next()
            if tok.kind == "eof":
                return
            if tok.kind == "punc":
                if tok.val in ("(", "[", "{", "<"):
                    depth += 1
                elif tok.val in (")", "]", "}", ">"):
                    depth -= 1
                elif tok.val == ">>":
                    depth -= 2
                elif tok.val == ";" and depth <= 0:
                    return
            nxt = self.peek()
            if depth <= 0 and nxt.nl and not (nxt.kind == "punc" and nxt.val in ("|", "&", ".", "<", "=>", "?", ":")) and \
                    not (tok.kind == "punc" and tok.val in ("|", "&", "=", "<", "=>", "?", ":", ",")):
                return

    def block(self):
        self.expect("{")
        body = []
        while not self.at("}"):
            if self.peek().kind == "eof":
                self.fail("unterminated block")
            body.append(self.statement())
        self.i += 1
        return body

    def import_decl(self):
        self.i += 1
        if self.peek().kind == "str":  # side-effect import
            mod = self.next().val
            self.semi()
            return ("import", [], mod, None, None)
        if self.at_id("type") and not self.at(",", 1) and not self.at_id("from", 1):
            while self.peek().kind != "str":
                self.i += 1
            self.i += 1
            self.semi()
            return ("empty",)
        default = star = None
        specs = []
        if self.peek().kind == "id":
            default = self.ident()
            self.eat(",")
        if self.eat("*"):
            if not self.eat_id("as"):
                self.fail("expected as")
            star = self.ident()
        elif self.eat("{"):
            while not self.eat("}"):
                is_type = self.at_id("type") and self.peek(1).kind == "id" and not self.at_id("as", 1)
                if is_type:
                    self.i += 1
                name = self.ident()
                alias = self.ident() if self.eat_id("as") else name
                if not is_type:
                    specs.append((name, alias))
                self.eat(",")
        if not self.eat_id("from"):
            self.fail("expected from")
        mod = self.next().val
        self.semi()
        return ("import", specs, mod, default, star)

    def export_decl(self):
        self.i += 1
        if self.at_id("type") and (self.at("{", 1) or self.at("*", 1)):  # export type { ... } [from ...]
            while not (self.at(";") or self.peek().nl and self.peek().kind != "str" and not self.at_id("from")):
                self.i += 1
                if self.peek().kind == "eof":
                    break
            self.eat(";")
            return ("empty",)
        if self.eat("*"):
            alias = self.ident() if self.eat_id("as") else None
            if not self.eat_id("from"):
                self.fail("expected from")
            mod = self.next().val
            self.semi()
            return ("export_star", mod, alias)
        if self.eat("{"):
            specs = []
            while not self.eat("}"):
                is_type = self.at_id("type") and self.peek(1).kind == "id" and not self.at_id("as", 1)
                if is_type:
                    self.i += 1
                name = self.ident()
                alias = self.ident() if self.eat_id("as") else name
                if not is_type:
                    specs.append((name, alias))
                self.eat(",")
            mod = None
            if self.eat_id("from"):
                mod = self.next().val
            self.semi()
            return ("export_names", specs, mod)
        if self.eat_id("default"):
            if self.at_id("function") or self.at_id("class"):
                d = self.statement()
                return ("export_default_decl", d)
            e = self.assignment()
            self.semi()
            return ("export_default", e)
        d = self.statement()
        return ("export", d)

    def var_decl(self, no_in=False):
        kind = self.next().val
        decls = []
        while True:
            pat = self.binding_pattern()
            if self.eat("!"):
                pass
            self.opt_type_annotation()
            init = self.assignment(no_in) if self.eat("=") else None
            decls.append((pat, init))
            if not self.eat(","):
                break
        return ("var", kind, decls)

    def binding_pattern(self):
        if self.eat("{"):
            props, rest = [], None
            while not self.eat("}"):
                if self.eat("..."):
                    rest = self.ident()
                else:
                    if self.at("["):
                        self.fail("computed keys in patterns are not supported")
                    tok = self.next()
                    key = tok.val if tok.kind in ("id", "str") else num_to_str(tok.val)
                    target = ("pid", key)
                    if self.eat(":"):
                        target = self.binding_pattern()
                    default = self.assignment() if self.eat("=") else None
                    props.append((key, target, default))
                self.eat(",")
            return ("pobj", props, rest)
        if self.eat("["):
            elems, rest = [], None
            while not self.eat("]"):
                if self.at(","):
                    self.i += 1
                    elems.append(None)
                    continue
                if self.eat("..."):
                    rest = self.binding_pattern()
                else:
                    target = self.binding_pattern()
                    default = self.assignment() if self.eat("=") else None
                    elems.append((target, default))
                self.eat(",")
            return ("parr", elems, rest)
        return ("pid", self.ident())

    def params(self):
        self.expect("(")
        out, props = [], []
        while not self.eat(")"):
            is_prop = False
            while self.peek().kind == "id" and self.peek().val in MODIFIERS and (self.peek(1).kind == "id" or self.at("{", 1) or self.at("[", 1)):
                is_prop = True
                self.i += 1
            rest = self.eat("...")
            if self.at_id("this") and self.at(":", 1):  # `this` parameter: type only
                self.i += 1
                self.opt_type_annotation()
                self.eat(",")
                continue
            pat = self.binding_pattern()
            self.eat("?")
            self.opt_type_annotation()
            default = self.assignment() if self.eat("=") else None
            out.append((pat, default, rest))
            if is_prop and pat[0] == "pid":
                props.append(pat[1])
            self.eat(",")
        return out, props

    def function_rest(self, name, is_arrow=False):
        if self.at("<"):
            self.skip_type_args()
        params, props = self.params()
        self.opt_type_annotation()
        if not self.at("{"):  # overload signature / abstract method: no body
            self.semi()
            return None
        body = self.block()
        return ("fn", params, body, False, False, name, props)

    def function_decl(self):
        self.i += 1
        if self.eat("*"):
            self.fail("generators are not supported")
        name = self.ident()
        fn = self.function_rest(name)
        if fn is None:
            return ("empty",)
        return ("fdecl", name, fn)

    def class_decl(self):
        self.i += 1
        name = self.ident() if self.peek().kind == "id" and not self.at_id("implements") and not self.at_id("extends") else None
        if self.at("<"):
            self.skip_type_args()
        parent = None
        if self.eat_id("extends"):
            parent = self.member_only()
            if self.at("<"):
                self.skip_type_args()
        if self.eat_id("implements"):
            while not self.at("{"):
                self.i += 1
        self.expect("{")
        members = []
        while not self.eat("}"):
            if self.eat(";"):
                continue
            static = False
            while self.peek().kind == "id" and self.peek().val in MODIFIERS and not (self.at("(", 1) or self.at("=", 1) or self.at(":", 1)
                                                                                      or self.at(";", 1) or self.at("<", 1) or self.at("?", 1)):
                if self.next().val == "static":
                    static = True
            if (self.at_id("get") or self.at_id("set")) and self.peek(1).kind in ("id", "str") and self.at("(", 2):
                self.fail("getters / setters are not supported")
            if self.at("["):
                self.fail("computed / index-signature members are not supported")
            tok = self.next()
            if tok.kind not in ("id", "str", "num"):
                self.i -= 1
                self.fail("class member expected")
            key = tok.val if tok.kind != "num" else num_to_str(tok.val)
            if self.at("(") or self.at("<"):
                fn = self.function_rest(key)
                if fn is not None:
                    members.append(("method", key, fn, static))
                continue
            self.eat("?")
            self.eat("!")
            self.opt_type_annotation()
            init = self.assignment() if self.eat("=") else None
            self.semi()
            members.append(("field", key, init, static))
        return ("cdecl", name, members, parent)

    def enum_decl(self):
        self.i += 1
        name = self.ident()
        self.expect("{")
        members = []
        while not self.eat("}"):
            tok = self.next()
            key = tok.val
            init = self.assignment() if self.eat("=") else None
            members.append((key, init))
            self.eat(",")
        return ("enum", name, members)

    def for_stmt(self):
        self.i += 1
        self.expect("(")
        init = None
        if self.at_id("const") or self.at_id("let") or self.at_id("var"):
            save = self.i
            kind = self.next().val
            pat = self.binding_pattern()
            self.opt_type_annotation()
            if self.eat_id("of"):
                it = self.assignment()
                self.expect(")")
                return ("forof", kind, pat, it, self.statement())
            if self.eat_id("in"):
                it = self.expression()
                self.expect(")")
                return ("forin", kind, pat, it, self.statement())
            self.i = save
            init = self.var_decl(no_in=True)
        elif not self.at(";"):
            init = ("expr", self.expression(no_in=True))
        self.expect(";")
        test = None if self.at(";") else self.expression()
        self.expect(";")
        upd = None if self.at(")") else self.expression()
        self.expect(")")
        return ("for", init, test, upd, self.statement())

    def switch_stmt(self):
        self.i += 1
        self.expect("(")
        disc = self.expression()
        self.expect(")")
        self.expect("{")
        cases = []
        while not self.eat("}"):
            if self.eat_id("default"):
                test = None
            elif self.eat_id("case"):
                test = self.expression()
            else:
                self.fail("case expected")
            self.expect(":")
            body = []
            while not (self.at_id("case") or self.at_id("default") or self.at("}")):
                body.append(self.statement())
            cases.append((test, body))
        return ("switch", disc, cases)

    def try_stmt(self):
        self.i += 1
        blk = self.block()
        param = handler = fin = None
        if self.eat_id("catch"):
            if self.eat("("):
                param = self.binding_pattern()
                self.opt_type_annotation()
                self.expect(")")
            handler = self.block()
        if self.eat_id("finally"):
            fin = self.block()
        return ("try", blk, param, handler, fin)

    # -- expressions
    def expression(self, no_in=False):
        e = self.assignment(no_in)
        if self.at(","):
            seq = [e]
            while self.eat(","):
                seq.append(self.assignment(no_in))
            return ("seq", seq)
        return e

    def is_arrow_ahead(self):
        """At '(' — is this the parameter list of an arrow function?"""
        j, depth = self.i, 0
        while True:
            tok = self.t[j]
            if tok.kind == "eof":
                return False
            if tok.kind == "punc":
                if tok.val in ("(", "[", "{"):
                    depth += 1
                elif tok.val in (")", "]", "}"):
                    depth -= 1
                    if depth == 0:
                        break
            j += 1
        nxt = self.t[j + 1]
        if nxt.kind == "punc" and nxt.val == "=>":
            return True
        if nxt.kind == "punc" and nxt.val == ":":  # (a: T): R => ...   — try to skip the return type
            save = self.i
            self.i = j + 2
            try:
                self.skip_type()
                ok = self.at("=>")
            except SyntaxError:
                ok = False
            self.i = save
            return ok
        return False

    def assignment(self, no_in=False):
        tok = self.peek()
        if tok.kind == "id" and tok.val not in RESERVED_STARTS and self.at("=>", 1):
            name = self.ident()
            self.i += 1
            return self.arrow_body([(("pid", name), None, False)])
        if tok.kind == "id" and tok.val == "async" and (self.at("(", 1) or self.at_id("function", 1)):
            self.fail("async functions are not supported")
        if self.at("(") and self.is_arrow_ahead():
            params, _ = self.params()
            self.opt_type_annotation()
            self.expect("=>")
            return self.arrow_body(params)
        if self.at("<") and self.peek(1).kind == "id":  # <T>(x: T) => ...
            save = self.i
            if self.skip_type_args() and self.at("(") and self.is_arrow_ahead():
                params, _ = self.params()
                self.opt_type_annotation()
                self.expect("=>")
                return self.arrow_body(params)
            self.i = save
        left = self.conditional(no_in)
        tok = self.peek()
        if tok.kind == "punc" and tok.val in ASSIGN_OPS:
            self.i += 1
            right = self.assignment(no_in)
            target = self.to_target(left)
            return ("assign", tok.val, target, right)
        return left

    def to_target(self, e):
        k = e[0]
        if k in ("id", "mem", "idx"):
            return e
        if k == "arr":
            elems, rest = [], None
            for el in e[1]:
                if el is None:
                    elems.append(None)
                elif el[0] == "spread":
                    rest = self.expr_to_pattern(el[1])
                else:
                    elems.append((self.expr_to_pattern(el), None))
            return ("parr", elems, rest)
        if k == "obj":
            props = []
            for p in e[1]:
                if p[0] != "prop" or p[1][0] != "str":
                    self.fail("unsupported destructuring assignment")
                props.append((p[1][1], self.expr_to_pattern(p[2]), None))
            return ("pobj", props, None)
        self.fail("invalid assignment target")

    def expr_to_pattern(self, e):
        if e[0] == "id":
            return ("pid", e[1])
        if e[0] in ("mem", "idx"):
            return ("ptarget", e)
        if e[0] == "assign" and e[1] == "=":
            return self.expr_to_pattern(e[2])
        return self.to_target(e)

    def arrow_body(self, params):
        if self.at("{"):
            return ("fn", params, self.block(), True, False, None, [])
        return ("fn", params, self.assignment(), True, True, None, [])

    def conditional(self, no_in=False):
        c = self.binary(0, no_in)
        if self.eat("?"):
            a = self.assignment()
            self.expect(":")
            b = self.assignment(no_in)
            return ("cond", c, a, b)
        return c

    def binary(self, min_prec, no_in=False):
        left = self.unary()
        while True:
            tok = self.peek()
            if tok.kind == "id" and tok.val in ("as", "satisfies") and not tok.nl:
                self.i += 1
                if self.eat_id("const"):
                    continue
                self.skip_type()
                continue
            op = tok.val if tok.kind == "punc" or (tok.kind == "id" and tok.val in ("instanceof", "in")) else None
            if op == "in" and no_in:
                break
            prec = BIN_PREC.get(op) if op is not None else None
            if prec is None or prec < min_prec:
                break
            self.i += 1
            right = self.binary(prec if op == "**" else prec + 1, no_in)
            left = ("log", op, left, right) if op in ("&&", "||", "??") else ("bin", op, left, right)
        return left

    def unary(self):
        tok = self.peek()
        if tok.kind == "punc":
            v = tok.val
            if v in ("!", "~", "+", "-"):
                self.i += 1
                return ("un", v, self.unary())
            if v in ("++", "--"):
                self.i += 1
                return ("upd", v, True, self.unary())
            if v == "<" and self.peek(1).kind == "id":  # <T>expr type assertion (old syntax)
                save = self.i
                if self.skip_type_args():      # (a generic arrow function was already recognised in assignment())
                    return self.unary()
                self.i = save
        elif tok.kind == "id" and tok.val in ("typeof", "void", "delete"):
            self.i += 1
            return ("un", tok.val, self.unary())
        elif tok.kind == "id" and tok.val == "await":
            self.fail("await is not supported")
        e = self.postfix()
        if self.at("**"):
            self.i += 1
            return ("bin", "**", e, self.unary())
        return e

    def postfix(self):
        e = self.call_member()
        tok = self.peek()
        if tok.kind == "punc" and tok.val in ("++", "--") and not tok.nl:
            self.i += 1
            return ("upd", tok.val, False, e)
        return e

    def arguments(self):
        self.expect("(")
        args = []
        while not self.eat(")"):
            if self.eat("..."):
                args.append(("spread", self.assignment()))
            else:
                args.append(self.assignment())
            self.eat(",")
        return args

    def call_member(self):
        if self.at_id("new"):
            self.i += 1
            if self.at("."):
                self.fail("new.target is not supported")
            callee = self.member_only()
            if self.at("<"):
                self.skip_type_args()
            args = self.arguments() if self.at("(") else []
            e = ("new", callee, args)
        else:
            e = self.primary()
        while True:
            tok = self.peek()
            if tok.kind == "punc":
                v = tok.val
                if v == ".":
                    self.i += 1
                    e = ("mem", e, self.next().val, False)
                    continue
                if v == "?.":
                    self.i += 1
                    if self.at("("):
                        e = ("call", e, self.arguments(), True)
                    elif self.eat("["):
                        idx = self.expression()
                        self.expect("]")
                        e = ("idx", e, idx, True)
                    else:
                        e = ("mem", e, self.next().val, True)
                    continue
                if v == "[" :
                    self.i += 1
                    idx = self.expression()
                    self.expect("]")
                    e = ("idx", e, idx, False)
                    continue
                if v == "(":
                    e = ("call", e, self.arguments(), False)
                    continue
                if v == "!" and not tok.nl:  # non-null assertion
                    self.i += 1
                    continue
                if v == "<" and e[0] in ("id", "mem"):  # f<T>(...)
                    save = self.i
                    if self.skip_type_args() and self.at("("):
                        continue
                    self.i = save
            elif tok.kind == "tpl" and not tok.nl and e[0] in ("id", "mem"):
                self.fail("tagged templates are not supported")
            break
        return e

    def member_only(self):
        e = self.primary()
        while True:
            if self.at("."):
                self.i += 1
                e = ("mem", e, self.next().val, False)
            elif self.at("["):
                self.i += 1
                idx = self.expression()
                self.expect("]")
                e = ("idx", e, idx, False)
            else:
                return e

    def primary(self):
        tok = self.next()
        k = tok.kind
        if k == "num":
            return ("num", tok.val)
        if k == "str":
            return ("str", tok.val)
        if k == "tpl":
            parts = []
            for p in tok.val:
                if isinstance(p, str):
                    parts.append(("str", p))
                else:
                    sub = Parser(p, self.fname)
                    parts.append(sub.expression())
            return ("tpl", parts)
        if k == "regex":
            return ("regex", tok.val[0], tok.val[1])
        if k == "id":
            v = tok.val
            if v == "this":
                return ("id", "this")
            if v == "null":
                return ("lit", None)
            if v == "undefined":
                return ("lit", UNDEF)
            if v == "true":
                return ("lit", True)
            if v == "false":
                return ("lit", False)
            if v == "function":
                if self.eat("*"):
                    self.fail("generators are not supported")
                name = self.ident() if self.peek().kind == "id" else None
                return self.function_rest(name)
            if v == "class":
                self.i -= 1
                d = self.class_decl()
                return ("class", d[1], d[2], d[3])
            if v == "super":
                return ("super",)
            return ("id", v)
        if k == "punc":
            v = tok.val
            if v == "(":
                e = self.expression()
                self.expect(")")
                return e
            if v == "[":
                elems = []
                while not self.eat("]"):
                    if self.at(","):
                        self.i += 1
                        elems.append(None)
                        continue
                    if self.eat("..."):
                        elems.append(("spread", self.assignment()))
                    else:
                        elems.append(self.assignment())
                    self.eat(",")
                return ("arr", elems)
            if v == "{":
                props = []
                while not self.eat("}"):
                    if self.eat("..."):
                        props.append(("spread", self.assignment()))
                        self.eat(",")
                        continue
                    if self.eat("["):
                        key = self.assignment()
                        self.expect("]")
                    else:
                        kt = self.next()
                        if (kt.kind == "id" and kt.val in ("get", "set", "async") and not (self.at(":") or self.at("(") or self.at(",") or self.at("}"))):
                            self.fail("accessors / async methods in object literals are not supported")
                        key = ("str", kt.val if kt.kind in ("id", "str") else num_to_str(kt.val))
                        if kt.kind == "id" and (self.at(",") or self.at("}")):  # shorthand
                            props.append(("prop", key, ("id", kt.val)))
                            self.eat(",")
                            continue
                    if self.at("(") or self.at("<"):
                        fn = self.function_rest(key[1] if key[0] == "str" else None)
                        props.append(("prop", key, fn))
                    else:
                        self.expect(":")
                        props.append(("prop", key, self.assignment()))
                    self.eat(",")
                return ("obj", props)
        self.i -= 1
        self.fail("expression expected")


# ---------------------------------------------------------------------------------------------------------------
# runtime environment
# ---------------------------------------------------------------------------------------------------------------
class Env:
    __slots__ = ("vars", "parent")

    def __init__(self, parent):
        self.vars = {}
        self.parent = parent


BREAK = ("b",)
CONT = ("c",)


class Interp:
    def __init__(self, log=None, stub_modules=(), virtual_modules=None, require=None, path_overrides=None):
        self.modules = {}
        self.require = require                        # host hook behind the global require(spec) (e.g. a fake native addon)
        self.path_overrides = dict(path_overrides or {})   # module location (as resolved) -> the file whose text is loaded there
        self.virtual_modules = dict(virtual_modules or {})   # bare specifier -> exports dict provided by the host (e.g. 'vitest')
        self.stub_modules = tuple(stub_modules)   # path suffixes loaded as EMPTY modules (e.g. the async WASM bridge)
        self.stubbed = []
        self.log = log if log is not None else (lambda *a: None)
        self.globals = Env(None)
        self.error_classes = {}
        self._install_globals()

    # -- host entry points -------------------------------------------------------------------------------------
    def load(self, path):
        if path in self.modules:
            return self.modules[path]
        path = os.path.realpath(path)
        if path in self.modules:
            return self.modules[path]
        exports = {}
        self.modules[path] = exports
        if any(path.endswith(sfx) for sfx in self.stub_modules):
            self.stubbed.append(path)
            return exports
        src = open(self.path_overrides.get(path, path), encoding="utf-8").read()
        ast = Parser(lex(src, path), path).program()
        env = Env(self.globals)
        env.vars["this"] = UNDEF
        Compiler(self, path, exports).run_module(ast, env)
        return exports

    def resolve(self, base, spec):
        if spec in self.virtual_modules:
            key = "virtual:" + spec
            self.modules[key] = self.virtual_modules[spec]
            return key
        if not spec.startswith("."):
            raise ImportError(f"{base}: bare module specifier {spec!r} is not supported")
        p = os.path.normpath(os.path.join(os.path.dirname(base), spec))
        for cand in (p, p + ".ts", p + ".js", os.path.join(p, "index.ts")):
            if cand in self.path_overrides or os.path.isfile(cand):
                return cand      # (an overridden module keeps its VIRTUAL location: its own imports resolve from there)
        if p.endswith(".js") and os.path.isfile(p[:-3] + ".ts"):
            return p[:-3] + ".ts"
        raise ImportError(f"{base}: cannot resolve {spec!r}")

    def call(self, f, this=UNDEF, args=()):
        return call_function(f, this, list(args))

    def construct(self, cls, args=()):
        return construct(cls, list(args))

    def get(self, obj, name):
        return get_prop(obj, name)

    def float32(self, data):
        return TypedArray("Float32Array", array("f", data))

    # -- globals -------------------------------------------------------------------------------------------------
    def _install_globals(self):
        g = self.globals.vars
        g["undefined"] = UNDEF
        g["NaN"] = math.nan
        g["Infinity"] = math.inf
        g["globalThis"] = g
        g["process"] = {"env": {}, "argv": []}

        def _require(spec=UNDEF):
            if self.require is None:
                throw_type_error("require is not available")
            return self.require(to_str(spec))
        g["require"] = _require
        M = {"PI": math.pi, "E": math.e, "LN2": math.log(2), "LN10": math.log(10), "LOG2E": 1 / math.log(2), "LOG10E": 1 / math.log(10),
             "SQRT2": math.sqrt(2), "SQRT1_2": math.sqrt(0.5),
             "abs": lambda x=UNDEF: abs(to_number(x)), "floor": lambda x=UNDEF: _floor(to_number(x)),
             "ceil": lambda x=UNDEF: _ceil(to_number(x)), "trunc": lambda x=UNDEF: _trunc(to_number(x)), "round": math_round,
             "sqrt": lambda x=UNDEF: _sqrt(to_number(x)), "max": math_max, "min": math_min,
             "pow": lambda a=UNDEF, b=UNDEF: js_pow(to_number(a), to_number(b)), "exp": _m1(math.exp), "log": _log(math.log),
             "log2": _log(math.log2), "log10": _log(math.log10), "sin": _m1(math.sin), "cos": _m1(math.cos), "tan": _m1(math.tan),
             "atan": _m1(math.atan), "asin": _m1(math.asin), "acos": _m1(math.acos), "tanh": _m1(math.tanh), "sinh": _m1(math.sinh),
             "cosh": _m1(math.cosh), "atan2": lambda a=UNDEF, b=UNDEF: math.atan2(to_number(a), to_number(b)), "sign": _sign,
             "fround": _fround, "hypot": _hypot, "cbrt": lambda x=UNDEF: math.copysign(abs(to_number(x)) ** (1.0 / 3.0), to_number(x)),
             "imul": lambda a=UNDEF, b=UNDEF: float(to_int32(float((to_int32(a) * to_int32(b)) & 0xFFFFFFFF))),
             "clz32": lambda x=UNDEF: float(32 - to_uint32(x).bit_length()),
             "random": lambda: (_ for _ in ()).throw(RuntimeError("Math.random() reached: golden runs must be deterministic"))}
        g["Math"] = Native("Math", props=M)
        N = {"MAX_VALUE": sys.float_info.max, "MIN_VALUE": 5e-324, "EPSILON": sys.float_info.epsilon, "MAX_SAFE_INTEGER": 9007199254740991.0,
             "MIN_SAFE_INTEGER": -9007199254740991.0, "POSITIVE_INFINITY": math.inf, "NEGATIVE_INFINITY": -math.inf, "NaN": math.nan,
             "isFinite": lambda x=UNDEF: type(x) is float and x == x and x not in (math.inf, -math.inf),
             "isNaN": lambda x=UNDEF: type(x) is float and x != x,
             "isInteger": lambda x=UNDEF: type(x) is float and x == x and x not in (math.inf, -math.inf) and x == math.floor(x),
             "isSafeInteger": lambda x=UNDEF: type(x) is float and x == x and abs(x) <= 9007199254740991.0 and x == math.floor(x),
             "parseFloat": lambda s=UNDEF: _parse_float(s), "parseInt": lambda s=UNDEF, r=UNDEF: _parse_int(s, r)}
        g["Number"] = Native("Number", call=lambda x=0.0: to_number(x), props=N)
        g["isNaN"] = lambda x=UNDEF: to_number(x) != to_number(x)
        g["isFinite"] = lambda x=UNDEF: (lambda v: v == v and v not in (math.inf, -math.inf))(to_number(x))
        g["parseFloat"] = N["parseFloat"]
        g["parseInt"] = N["parseInt"]
        g["String"] = Native("String", call=lambda x="": to_str(x))
        g["Boolean"] = Native("Boolean", call=lambda x=False: truthy(x))
        g["Symbol"] = Native("Symbol", props={"iterator": "@@iterator"})
        out = lambda *a: self.log(" ".join(to_str(x) if not isinstance(x, (dict, list)) else json_stringify(x) for x in a))
        g["console"] = Native("console", props={k: out for k in ("log", "warn", "error", "info", "debug", "time", "timeEnd", "table")})
        clock = [0.0]

        def now():
            clock[0] += 1.0
            return clock[0]
        g["performance"] = Native("performance", props={"now": now})
        g["Date"] = Native("Date", props={"now": now})
        g["JSON"] = Native("JSON", props={"stringify": lambda v=UNDEF, *_: json_stringify(v)})
        g["Object"] = Native("Object", call=lambda v=UNDEF: {} if v is UNDEF or v is None else v, props={
            "keys": lambda o: [k for k in own_keys(o)], "values": lambda o: [get_prop(o, k) for k in own_keys(o)],
            "entries": lambda o: [[k, get_prop(o, k)] for k in own_keys(o)],
            "assign": _object_assign, "freeze": lambda o: o, "isFrozen": lambda o: False, "create": lambda p=None, *_: {},
            "fromEntries": lambda it: {prop_key(e[0]): e[1] for e in iterate(it)}}, construct=lambda args: {})
        g["Array"] = Native("Array", call=lambda *a: _array_ctor(list(a)), construct=_array_ctor,
                            props={"from": _array_from, "isArray": lambda v=UNDEF: type(v) is list, "of": lambda *a: list(a)})
        for name, code in KIND_CODE.items():
            g[name] = Native(name, construct=_typed_ctor(name, code),
                             props={"BYTES_PER_ELEMENT": float(array(code).itemsize), "from": (lambda nm, cd: lambda src=UNDEF, fn=UNDEF:
                                    _typed_ctor(nm, cd)([_array_from(src, fn)]))(name, code)})
        g["Map"] = Native("Map", construct=_map_ctor)
        g["Set"] = Native("Set", construct=_set_ctor)
        g["WeakMap"] = Native("WeakMap", construct=_map_ctor)   # (object keys compare by identity in map_key: same behaviour)
        g["WeakSet"] = Native("WeakSet", construct=_set_ctor)
        for name in ("Error", "TypeError", "RangeError", "SyntaxError", "ReferenceError", "EvalError"):
            cls = JSClass(name)
            cls.native_new = (lambda c: lambda args: JSObj(c, {"message": to_str(args[0]) if args and args[0] is not UNDEF else "",
                                                               "name": c.name, "stack": ""}))(cls)
            self.error_classes[name] = cls
            g[name] = cls
        ERR.update(self.error_classes)


ERR = {}


def throw_type_error(msg):
    cls = ERR.get("TypeError")
    raise JSThrow(JSObj(cls, {"message": msg, "name": "TypeError", "stack": ""}))


def _parse_float(s):
    m = re.match(r"\s*([+-]?(?:Infinity|\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?))", to_str(s))
    return float(m.group(1).replace("Infinity", "inf")) if m else math.nan


def _parse_int(s, radix=UNDEF):
    r = int(to_number(radix)) if radix is not UNDEF and to_number(radix) == to_number(radix) else 10
    m = re.match(r"\s*([+-]?)(0[xX])?([0-9a-zA-Z]*)", to_str(s))
    sign, hexp, digits = m.group(1), m.group(2), m.group(3)
    if hexp and r in (10, 16):
        r = 16
    elif hexp:
        digits = "0"
    ok = ""
    for ch in digits:
        d = int(ch, 36)
        if d >= r:
            break
        ok += ch
    if not ok:
        return math.nan
    v = float(int(ok, r))
    return -v if sign == "-" else v


def _object_assign(target, *sources):
    for s in sources:
        if s is None or s is UNDEF:
            continue
        for k in own_keys(s):
            set_prop(target, k, get_prop(s, k))
    return target


def own_keys(o):
    if type(o) is dict:
        return list(o.keys())
    if type(o) is JSObj:
        return list(o.props.keys())
    if type(o) is list:
        return [str(i) for i in range(len(o))]
    if type(o) is TypedArray:
        return [str(i) for i in range(len(o.a))]
    if type(o) is str:
        return [str(i) for i in range(len(o))]
    return []


def json_stringify(v):
    t = type(v)
    if v is None:
        return "null"
    if v is UNDEF or t in (JSFunction, Native, BoundMethod, JSClass):
        return UNDEF
    if t is bool:
        return "true" if v else "false"
    if t is float:
        return num_to_str(v) if (v == v and v not in (math.inf, -math.inf)) else "null"
    if t is str:
        import json
        return json.dumps(v, ensure_ascii=False)
    if t is list:
        return "[" + ",".join((lambda s: "null" if s is UNDEF else s)(json_stringify(e)) for e in v) + "]"
    if t is TypedArray:
        return "{" + ",".join(f'"{i}":{json_stringify(float(e))}' for i, e in enumerate(v.a)) + "}"
    if t is JSMap or t is JSSet:
        return "{}"
    import json
    items = []
    for k in own_keys(v):
        s = json_stringify(get_prop(v, k))
        if s is not UNDEF:
            items.append(json.dumps(k, ensure_ascii=False) + ":" + s)
    return "{" + ",".join(items) + "}"


def iterate(v):
    t = type(v)
    if t is list:
        return list(v)
    if t is TypedArray:
        return [float(x) for x in v.a]
    if t is str:
        return list(v)
    if t is JSMap:
        return [[k, val] for k, val in v.d.values()]
    if t is JSSet:
        return list(v.d.values())
    if t is dict and "length" in v:  # array-like
        return [v.get(str(i), UNDEF) for i in range(int(to_number(v["length"])))]
    throw_type_error("object is not iterable")


def _array_ctor(args):
    if len(args) == 1 and type(args[0]) is float:
        return [UNDEF] * int(args[0])
    return list(args)


def _array_from(src=UNDEF, fn=UNDEF, *_):
    if type(src) is dict and "length" in src and not isinstance(src.get("length"), (list, dict)):
        n = int(to_number(src["length"]))
        items = [src.get(str(i), UNDEF) for i in range(n)]
    else:
        items = iterate(src)
    if fn is UNDEF:
        return items
    return [call_function(fn, UNDEF, [x, float(i)]) for i, x in enumerate(items)]


def _typed_ctor(name, code):
    def make(args):
        a0 = args[0] if args else UNDEF
        if a0 is UNDEF:
            return TypedArray(name, array(code))
        if type(a0) is float:
            return TypedArray(name, array(code, bytes(array(code).itemsize * int(a0))))
        if type(a0) is TypedArray and a0.a.typecode == code and len(args) == 1:
            return TypedArray(name, array(code, a0.a))
        if type(a0) is dict and a0.get("__arraybuffer__") is not None:
            raise NotImplementedError("ArrayBuffer views are not supported")
        out = TypedArray(name, array(code, bytes(array(code).itemsize * len(iterate(a0)))))
        for i, x in enumerate(iterate(a0)):
            out.store(i, x)
        return out
    return make


def _map_ctor(args):
    m = JSMap()
    if args and args[0] is not UNDEF and args[0] is not None:
        for e in iterate(args[0]):
            m.d[map_key(e[0])] = (e[0], e[1])
    return m


def _set_ctor(args):
    s = JSSet()
    if args and args[0] is not UNDEF and args[0] is not None:
        for e in iterate(args[0]):
            s.d[map_key(e)] = e
    return s


# ---------------------------------------------------------------------------------------------------------------
# property access, calls
# ---------------------------------------------------------------------------------------------------------------
def idx_int(k):
    if type(k) is float:
        i = int(k)
        return i if i == k and i >= 0 else -1
    if type(k) is str and k.isdigit():
        return int(k)
    return -1


def get_prop(obj, key):
    t = type(obj)
    if t is TypedArray:
        if type(key) is float:
            i = int(key)
            if i == key and 0 <= i < len(obj.a):
                return float(obj.a[i])
            return UNDEF
        if key == "length":
            return float(len(obj.a))
        i = idx_int(key)
        if i >= 0:
            return float(obj.a[i]) if i < len(obj.a) else UNDEF
        if key == "BYTES_PER_ELEMENT":
            return float(obj.a.itemsize)
        if key == "byteLength":
            return float(obj.a.itemsize * len(obj.a))
        if key in TYPED_METHODS:
            return BoundMethod(obj, TYPED_METHODS[key])
        return UNDEF
    if t is list:
        if type(key) is float:
            i = int(key)
            if i == key and 0 <= i < len(obj):
                return obj[i]
            return UNDEF
        if key == "length":
            return float(len(obj))
        i = idx_int(key)
        if i >= 0:
            return obj[i] if i < len(obj) else UNDEF
        if key in ARRAY_METHODS:
            return BoundMethod(obj, ARRAY_METHODS[key])
        return UNDEF
    if t is JSObj:
        k = key if type(key) is str else prop_key(key)
        p = obj.props
        if k in p:
            return p[k]
        cls = obj.cls
        if cls is not None:
            f = cls.find_method(k)
            if f is not None:
                return BoundMethod(obj, f)
        if k == "constructor":
            return cls
        if k == "toString":
            return BoundMethod(obj, lambda this: to_str(this))
        return UNDEF
    if t is dict:
        k = key if type(key) is str else prop_key(key)
        v = obj.get(k, UNDEF)
        if v is UNDEF and k == "hasOwnProperty":
            return BoundMethod(obj, lambda this, name=UNDEF: prop_key(name) in this)
        return v
    if t is str:
        if key == "length":
            return float(len(obj))
        i = idx_int(key)
        if i >= 0:
            return obj[i] if i < len(obj) else UNDEF
        if key in STRING_METHODS:
            return BoundMethod(obj, STRING_METHODS[key])
        return UNDEF
    if t is float:
        if key in NUMBER_METHODS:
            return BoundMethod(obj, NUMBER_METHODS[key])
        return UNDEF
    if t is Native:
        k = key if type(key) is str else prop_key(key)
        if k in obj.props:
            return obj.props[k]
        if k == "name":
            return obj.name
        return UNDEF
    if t is JSClass:
        k = key if type(key) is str else prop_key(key)
        v = obj.find_static(k)
        if v is UNDEF and k == "name":
            return obj.name
        return v
    if t is JSRegExp:
        if key == "source":
            return obj.source
        if key == "flags":
            return obj.flags
        if key == "exec":
            return BoundMethod(obj, _regexp_exec)
        if key == "test":
            return BoundMethod(obj, lambda this, s=UNDEF: this.rx.search(to_str(s)) is not None)
        return UNDEF
    if t is JSMap or t is JSSet:
        if key == "size":
            return float(len(obj.d))
        table = MAP_METHODS if t is JSMap else SET_METHODS
        if key in table:
            return BoundMethod(obj, table[key])
        return UNDEF
    if t is JSFunction or t is BoundMethod:
        if key == "name":
            return (obj.name if t is JSFunction else "") or ""
        if key == "length":
            return float(len(obj.params)) if t is JSFunction else 0.0
        if key in FUNCTION_METHODS:
            return BoundMethod(obj, FUNCTION_METHODS[key])
        return UNDEF
    if t is bool:
        if key == "toString":
            return BoundMethod(obj, lambda this: to_str(this))
        return UNDEF
    if obj is None or obj is UNDEF:
        throw_type_error(f"Cannot read properties of {to_str(obj)} (reading '{to_str(key)}')")
    if callable(obj):
        return UNDEF
    raise RuntimeError(f"get_prop on unsupported host value {type(obj)}")


def _regexp_exec(this, s=UNDEF):
    m = this.rx.search(to_str(s))
    if m is None:
        return None
    return [m.group(0)] + [UNDEF if g is None else g for g in m.groups()]


def set_prop(obj, key, val):
    t = type(obj)
    if t is TypedArray:
        if type(key) is float:
            i = int(key)
            if i == key and 0 <= i < len(obj.a):
                obj.store(i, val)
            return
        i = idx_int(key)
        if 0 <= i < len(obj.a):
            obj.store(i, val)
        return
    if t is list:
        if type(key) is float:
            i = int(key)
            if i == key and i >= 0:
                if i < len(obj):
                    obj[i] = val
                else:
                    obj.extend([UNDEF] * (i - len(obj)))
                    obj.append(val)
                return
        if key == "length":
            n = int(to_number(val))
            if n < len(obj):
                del obj[n:]
            else:
                obj.extend([UNDEF] * (n - len(obj)))
            return
        i = idx_int(key)
        if i >= 0:
            return set_prop(obj, float(i), val)
        throw_type_error("named properties on arrays are not supported")
    if t is JSObj:
        obj.props[key if type(key) is str else prop_key(key)] = val
        return
    if t is dict:
        obj[key if type(key) is str else prop_key(key)] = val
        return
    if t is JSClass:
        obj.statics[prop_key(key)] = val
        return
    if t is Native:
        obj.props[prop_key(key)] = val
        return
    if obj is None or obj is UNDEF:
        throw_type_error(f"Cannot set properties of {to_str(obj)} (setting '{to_str(key)}')")
    # primitives: silently ignored (sloppy) — strict mode would throw, the reference never does this


def call_function(f, this, args):
    t = type(f)
    if t is JSFunction:
        env = Env(f.env)
        v = env.vars
        if not f.is_arrow:
            v["this"] = this
        if f.simple is not None:
            names = f.simple
            n = len(args)
            for i, name in enumerate(names):
                v[name] = args[i] if i < n else UNDEF
        else:
            bind_params(f.params, args, env)
        if f.is_expr:
            return f.body(env)
        r = f.body(env)
        if r is None:
            return UNDEF
        return r[1]
    if t is BoundMethod:
        fn = f.fn
        if type(fn) is JSFunction:
            return call_function(fn, f.this, args)
        return fn(f.this, *args)
    if t is Native:
        if f.call is None:
            throw_type_error(f"{f.name} is not a function")
        return f.call(*args)
    if t is JSClass:
        throw_type_error(f"Class constructor {f.name} cannot be invoked without 'new'")
    if callable(f):
        return f(*args)
    throw_type_error(f"{to_str(f)} is not a function")


def construct(cls, args):
    t = type(cls)
    if t is JSClass:
        if cls.native_new is not None:
            return cls.native_new(args)
        obj = JSObj(cls)
        r = init_instance(cls, obj, args)
        return r if r is not None else obj
    if t is Native and cls.construct is not None:
        return cls.construct(args)
    if t is JSFunction:
        obj = JSObj(None)
        r = call_function(cls, obj, args)
        return r if isinstance(r, (JSObj, dict, list)) else obj
    throw_type_error(f"{to_str(cls)} is not a constructor")


def init_instance(cls, obj, args):
    """Runs the construction steps of `cls` on obj — TypeScript's emit order: [super(...)], parameter properties, field
    initialisers, then the rest of the constructor body.  -> an object the constructor returned explicitly, or None."""
    ctor, parent = cls.ctor, cls.parent

    def own_fields():
        for name, init in cls.fields:
            obj.props[name] = init(obj) if init is not None else UNDEF
    if ctor is None:
        if parent is not None:      # implicit constructor(...args) { super(...args); }
            init_instance(parent, obj, args)
        own_fields()
        return None
    env = Env(ctor.env)
    env.vars["this"] = obj
    bind_params(ctor.params, args, env)

    def own_start():
        for name in ctor.param_props:
            obj.props[name] = env.vars[name]
        own_fields()
    if parent is None:
        own_start()
    else:
        called = [False]

        def super_ctor(*a):
            if called[0]:
                raise JSThrow(JSObj(ERR.get("ReferenceError"), {"message": "Super constructor may only be called once",
                                                                "name": "ReferenceError", "stack": ""}))
            called[0] = True
            init_instance(parent, obj, list(a))
            own_start()
            return UNDEF
        env.vars["%super_ctor"] = super_ctor
    r = ctor.body(env)
    if r is not None and isinstance(r[1], (JSObj, dict, list)):
        return r[1]
    return None


def bind_params(params, args, env):
    n = len(args)
    for i, (pat, default, rest) in enumerate(params):
        if rest:
            bind_pattern(pat, list(args[i:]), env)
            return
        v = args[i] if i < n else UNDEF
        if v is UNDEF and default is not None:
            v = default(env)
        bind_pattern(pat, v, env)


def bind_pattern(pat, v, env, declare=True):
    k = pat[0]
    if k == "pid":
        if declare:
            env.vars[pat[1]] = v
        else:
            assign_var(env, pat[1], v)
    elif k == "ptarget":
        pat[1](env, v)
    elif k == "pobj":
        if v is None or v is UNDEF:
            throw_type_error("Cannot destructure 'undefined' or 'null'")
        seen = []
        for key, target, default in pat[1]:
            x = get_prop(v, key)
            if x is UNDEF and default is not None:
                x = default(env)
            seen.append(key)
            bind_pattern(target, x, env, declare)
        if pat[2] is not None:
            rest = {kk: get_prop(v, kk) for kk in own_keys(v) if kk not in seen}
            bind_pattern(("pid", pat[2]), rest, env, declare)
    else:  # parr
        items = iterate(v)
        for i, el in enumerate(pat[1]):
            if el is None:
                continue
            x = items[i] if i < len(items) else UNDEF
            if x is UNDEF and el[1] is not None:
                x = el[1](env)
            bind_pattern(el[0], x, env, declare)
        if pat[2] is not None:
            bind_pattern(pat[2], items[len(pat[1]):], env, declare)


def assign_var(env, name, v):
    e = env
    while e is not None:
        if name in e.vars:
            e.vars[name] = v
            return
        e = e.parent
    raise JSThrow(JSObj(ERR.get("ReferenceError"), {"message": f"{name} is not defined", "name": "ReferenceError", "stack": ""}))


# -- methods of the built-in types ---------------------------------------------------------------------------------
def _cmp_default(a, b):
    if a is UNDEF:
        return 0 if b is UNDEF else 1
    if b is UNDEF:
        return -1
    sa, sb = to_str(a), to_str(b)
    return -1 if sa < sb else (1 if sa > sb else 0)


def _sort(this, fn=UNDEF):
    import functools
    if fn is UNDEF:
        key = functools.cmp_to_key(_cmp_default)
    else:
        def cmp(a, b):
            r = to_number(call_function(fn, UNDEF, [a, b]))
            return -1 if r < 0 else (1 if r > 0 else 0)
        key = functools.cmp_to_key(cmp)
    if type(this) is TypedArray:
        if fn is UNDEF:
            nan = [x for x in this.a if x != x]
            vals = sorted(x for x in this.a if x == x)
            this.a[:] = array(this.a.typecode, vals + nan)
        else:
            this.a[:] = array(this.a.typecode, sorted((float(x) for x in this.a), key=key))
        return this
    und = [x for x in this if x is UNDEF]
    rest = sorted((x for x in this if x is not UNDEF), key=key)   # list.sort is stable, like Array.prototype.sort
    this[:] = rest + und
    return this


def _norm_index(v, n, default):
    if v is UNDEF:
        return default
    x = to_number(v)
    if x != x:
        return 0
    x = math.trunc(x) if x not in (math.inf, -math.inf) else x
    if x < 0:
        return int(max(0, n + x))
    return int(min(x, n))


def _each(this):
    return [float(x) for x in this.a] if type(this) is TypedArray else this


def _arr_map(this, fn, thisArg=UNDEF):
    out = [call_function(fn, thisArg, [x, float(i), this]) for i, x in enumerate(_each(this))]
    if type(this) is TypedArray:
        res = TypedArray(this.kind, array(this.a.typecode, bytes(this.a.itemsize * len(out))))
        for i, x in enumerate(out):
            res.store(i, x)
        return res
    return out


def _arr_filter(this, fn, thisArg=UNDEF):
    out = [x for i, x in enumerate(_each(this)) if truthy(call_function(fn, thisArg, [x, float(i), this]))]
    return TypedArray(this.kind, array(this.a.typecode, out)) if type(this) is TypedArray else out


def _arr_reduce(this, fn, *init):
    items = _each(this)
    if init:
        acc, start = init[0], 0
    else:
        if not items:
            throw_type_error("Reduce of empty array with no initial value")
        acc, start = items[0], 1
    for i in range(start, len(items)):
        acc = call_function(fn, UNDEF, [acc, items[i], float(i), this])
    return acc


def _arr_foreach(this, fn, thisArg=UNDEF):
    for i, x in enumerate(_each(this)):
        call_function(fn, thisArg, [x, float(i), this])
    return UNDEF


def _arr_slice(this, a=UNDEF, b=UNDEF):
    n = len(this.a) if type(this) is TypedArray else len(this)
    lo, hi = _norm_index(a, n, 0), _norm_index(b, n, n)
    if type(this) is TypedArray:
        return TypedArray(this.kind, array(this.a.typecode, this.a[lo:hi]))
    return this[lo:hi]


def _arr_fill(this, v=UNDEF, a=UNDEF, b=UNDEF):
    n = len(this.a) if type(this) is TypedArray else len(this)
    lo, hi = _norm_index(a, n, 0), _norm_index(b, n, n)
    for i in range(lo, hi):
        if type(this) is TypedArray:
            this.store(i, v)
        else:
            this[i] = v
    return this


def _arr_indexof(this, v=UNDEF, start=UNDEF):
    items = _each(this)
    for i in range(_norm_index(start, len(items), 0), len(items)):
        if strict_eq(items[i], v):
            return float(i)
    return -1.0


def _arr_includes(this, v=UNDEF):
    return any(strict_eq(x, v) or (type(x) is float and type(v) is float and x != x and v != v) for x in _each(this))


def _arr_join(this, sep=UNDEF):
    s = "," if sep is UNDEF else to_str(sep)
    return s.join("" if (x is None or x is UNDEF) else to_str(x) for x in _each(this))


def _arr_find(this, fn):
    for i, x in enumerate(_each(this)):
        if truthy(call_function(fn, UNDEF, [x, float(i), this])):
            return x
    return UNDEF


def _arr_findindex(this, fn):
    for i, x in enumerate(_each(this)):
        if truthy(call_function(fn, UNDEF, [x, float(i), this])):
            return float(i)
    return -1.0


def _arr_splice(this, start=UNDEF, count=UNDEF, *items):
    n = len(this)
    lo = _norm_index(start, n, 0)
    cnt = n - lo if count is UNDEF else int(max(0, min(to_number(count), n - lo)))
    removed = this[lo:lo + cnt]
    this[lo:lo + cnt] = list(items)
    return removed


def _arr_concat(this, *others):
    out = list(this)
    for o in others:
        if type(o) is list:
            out.extend(o)
        else:
            out.append(o)
    return out


def _arr_flat(this, depth=UNDEF):
    d = 1 if depth is UNDEF else int(to_number(depth))
    out = []
    for x in this:
        if type(x) is list and d > 0:
            out.extend(_arr_flat(x, float(d - 1)))
        else:
            out.append(x)
    return out


def _arr_reverse(this):
    if type(this) is TypedArray:
        this.a.reverse()
    else:
        this.reverse()
    return this


def _arr_pop(this):
    return this.pop() if this else UNDEF


def _arr_push(this, *items):
    this.extend(items)
    return float(len(this))


def _arr_shift(this):
    return this.pop(0) if this else UNDEF


def _arr_unshift(this, *items):
    this[0:0] = list(items)
    return float(len(this))


ARRAY_METHODS = {
    "push": _arr_push, "pop": _arr_pop, "shift": _arr_shift, "unshift": _arr_unshift, "map": _arr_map, "filter": _arr_filter,
    "reduce": _arr_reduce, "forEach": _arr_foreach, "slice": _arr_slice, "sort": _sort, "reverse": _arr_reverse, "join": _arr_join,
    "fill": _arr_fill, "indexOf": _arr_indexof, "includes": _arr_includes, "concat": _arr_concat, "find": _arr_find,
    "findIndex": _arr_findindex, "splice": _arr_splice, "flat": _arr_flat,
    "every": lambda this, fn: all(truthy(call_function(fn, UNDEF, [x, float(i), this])) for i, x in enumerate(_each(this))),
    "some": lambda this, fn: any(truthy(call_function(fn, UNDEF, [x, float(i), this])) for i, x in enumerate(_each(this))),
    "keys": lambda this: [float(i) for i in range(len(_each(this)))], "values": lambda this: list(_each(this)),
    "entries": lambda this: [[float(i), x] for i, x in enumerate(_each(this))], "toString": lambda this: to_str(this),
    "at": lambda this, i=0.0: (lambda it, j: it[j] if -len(it) <= j < len(it) else UNDEF)(_each(this), int(to_number(i))),
    "flatMap": lambda this, fn: _arr_flat(_arr_map(this, fn)),
}


def _typed_set(this, src, offset=UNDEF):
    off = 0 if offset is UNDEF else int(to_number(offset))
    items = iterate(src)
    if off + len(items) > len(this.a):
        raise JSThrow(JSObj(ERR.get("RangeError"), {"message": "offset is out of bounds", "name": "RangeError", "stack": ""}))
    for i, x in enumerate(items):
        this.store(off + i, x)
    return UNDEF


def _typed_subarray(this, a=UNDEF, b=UNDEF):
    # a COPY, not a view on shared memory: enough for read-only uses (the drop-in class reads rows through it); code
    # that writes through a subarray would need real views — the reference's sources never call subarray
    return _arr_slice(this, a, b)


TYPED_METHODS = {k: ARRAY_METHODS[k] for k in ("map", "filter", "reduce", "forEach", "slice", "sort", "reverse", "join", "fill", "indexOf",
                                                "includes", "find", "findIndex", "every", "some", "keys", "values", "entries", "toString", "at")}
TYPED_METHODS.update({"set": _typed_set, "subarray": _typed_subarray})

STRING_METHODS = {
    "charAt": lambda s, i=0.0: s[int(to_number(i))] if 0 <= int(to_number(i)) < len(s) else "",
    "charCodeAt": lambda s, i=0.0: float(ord(s[int(to_number(i))])) if 0 <= int(to_number(i)) < len(s) else math.nan,
    "padStart": lambda s, n, c=" ": s.rjust(int(to_number(n)), to_str(c)[:1] or " "),
    "padEnd": lambda s, n, c=" ": s.ljust(int(to_number(n)), to_str(c)[:1] or " "),
    "toString": lambda s: s, "slice": lambda s, a=UNDEF, b=UNDEF: s[_norm_index(a, len(s), 0):_norm_index(b, len(s), len(s))],
    "substring": lambda s, a=UNDEF, b=UNDEF: s[_norm_index(a, len(s), 0):_norm_index(b, len(s), len(s))],
    "split": lambda s, sep=UNDEF: [s] if sep is UNDEF else (list(s) if to_str(sep) == "" else s.split(to_str(sep))),
    "repeat": lambda s, n=0.0: s * int(to_number(n)), "toUpperCase": lambda s: s.upper(), "toLowerCase": lambda s: s.lower(),
    "trim": lambda s: s.strip(), "includes": lambda s, x="": to_str(x) in s, "indexOf": lambda s, x="": float(s.find(to_str(x))),
    "startsWith": lambda s, x="": s.startswith(to_str(x)), "endsWith": lambda s, x="": s.endswith(to_str(x)),
    "concat": lambda s, *a: s + "".join(to_str(x) for x in a),
}


def _to_fixed(x, d=0.0):
    d = int(to_number(d))
    if x != x:
        return "NaN"
    if x in (math.inf, -math.inf):
        return num_to_str(x)
    from decimal import Decimal, ROUND_HALF_UP
    q = Decimal(1).scaleb(-d)
    return str(Decimal(x).quantize(q, rounding=ROUND_HALF_UP))


NUMBER_METHODS = {"toFixed": _to_fixed, "toString": lambda x, r=UNDEF: num_to_str(x) if r is UNDEF or to_number(r) == 10 else _radix(x, int(to_number(r))),
                  "toPrecision": lambda x, p=UNDEF: num_to_str(x) if p is UNDEF else f"{x:.{int(to_number(p))}g}", "valueOf": lambda x: x}


def _radix(x, r):
    n = int(x)
    digits = "0123456789abcdefghijklmnopqrstuvwxyz"
    if n == 0:
        return "0"
    s, m = "", abs(n)
    while m:
        s = digits[m % r] + s
        m //= r
    return ("-" if n < 0 else "") + s


MAP_METHODS = {
    "get": lambda m, k=UNDEF: m.d.get(map_key(k), (None, UNDEF))[1],
    "set": lambda m, k=UNDEF, v=UNDEF: (m.d.__setitem__(map_key(k), (k, v)), m)[1],
    "has": lambda m, k=UNDEF: map_key(k) in m.d, "delete": lambda m, k=UNDEF: m.d.pop(map_key(k), None) is not None,
    "clear": lambda m: (m.d.clear(), UNDEF)[1], "keys": lambda m: [k for k, _ in m.d.values()], "values": lambda m: [v for _, v in m.d.values()],
    "entries": lambda m: [[k, v] for k, v in m.d.values()],
    "forEach": lambda m, fn: ([call_function(fn, UNDEF, [v, k, m]) for k, v in list(m.d.values())], UNDEF)[1],
}
SET_METHODS = {
    "add": lambda s, v=UNDEF: (s.d.__setitem__(map_key(v), v), s)[1], "has": lambda s, v=UNDEF: map_key(v) in s.d,
    "delete": lambda s, v=UNDEF: s.d.pop(map_key(v), None) is not None, "clear": lambda s: (s.d.clear(), UNDEF)[1],
    "values": lambda s: list(s.d.values()), "keys": lambda s: list(s.d.values()),
    "forEach": lambda s, fn: ([call_function(fn, UNDEF, [v, v, s]) for v in list(s.d.values())], UNDEF)[1],
}
FUNCTION_METHODS = {
    "call": lambda f, this=UNDEF, *a: call_function(f, this, list(a)),
    "apply": lambda f, this=UNDEF, a=UNDEF: call_function(f, this, [] if a is UNDEF or a is None else iterate(a)),
    "bind": lambda f, this=UNDEF, *a: BoundMethod(this, (lambda fn, pre: lambda t, *rest: call_function(fn, t, list(pre) + list(rest)))(f, a)),
}


# ---------------------------------------------------------------------------------------------------------------
# compiler: AST -> Python closures  (expr closures: f(env) -> value; statement closures: f(env) -> None | BREAK | CONT | ("r", v))
# ---------------------------------------------------------------------------------------------------------------
class Compiler:
    def __init__(self, interp, path, exports):
        self.interp, self.path, self.exports = interp, path, exports

    # -- modules
    def run_module(self, ast, env):
        names = []          # (local name, exported name) resolved after the body has run
        for st in ast:      # function declarations are hoisted
            inner = st[1] if st[0] == "export" else st
            if inner[0] == "fdecl":
                env.vars[inner[1]] = self.make_function(inner[2])(env)
        for st in ast:
            k = st[0]
            if k == "import":
                _, specs, mod, default, star = st
                ex = self.interp.load(self.interp.resolve(self.path, mod))
                for name, alias in specs:
                    if name not in ex:
                        raise ImportError(f"{self.path}: {mod} has no export {name!r} (yet: circular import?)")
                    env.vars[alias] = ex[name]
                if default is not None:
                    env.vars[default] = ex.get("default", UNDEF)
                if star is not None:
                    env.vars[star] = dict(ex)
            elif k == "export_star":
                ex = self.interp.load(self.interp.resolve(self.path, st[1]))
                if st[2] is not None:
                    self.exports[st[2]] = dict(ex)
                else:
                    for name, v in ex.items():
                        if name != "default":
                            self.exports[name] = v
            elif k == "export_names":
                _, specs, mod = st
                if mod is not None:
                    ex = self.interp.load(self.interp.resolve(self.path, mod))
                    for name, alias in specs:
                        self.exports[alias] = ex[name]
                else:
                    names.extend(specs)
            elif k == "export_default":
                self.exports["default"] = self.expr(st[1])(env)
            elif k == "export_default_decl":
                d = st[1]
                self.stmt(d)(env)
                if d[0] in ("fdecl", "cdecl") and d[1]:
                    self.exports["default"] = env.vars[d[1]]
            elif k == "export":
                d = st[1]
                if d[0] != "fdecl":
                    self.stmt(d)(env)
                for name in self.declared_names(d):
                    self.exports[name] = env.vars[name]
            elif k == "fdecl":
                pass
            else:
                r = self.stmt(st)(env)
                if r is not None:
                    raise SyntaxError(f"{self.path}: illegal top-level completion")
        for name, alias in names:
            self.exports[alias] = env.vars[name]

    def declared_names(self, d):
        k = d[0]
        if k == "var":
            out = []
            for pat, _ in d[2]:
                self.pattern_names(pat, out)
            return out
        if k in ("fdecl", "cdecl", "enum"):
            return [d[1]]
        return []

    def pattern_names(self, pat, out):
        if pat[0] == "pid":
            out.append(pat[1])
        elif pat[0] == "pobj":
            for _, target, _ in pat[1]:
                self.pattern_names(target, out)
            if pat[2]:
                out.append(pat[2])
        elif pat[0] == "parr":
            for el in pat[1]:
                if el is not None:
                    self.pattern_names(el[0], out)
            if pat[2]:
                self.pattern_names(pat[2], out)

    # -- patterns: compile defaults and member targets once
    def pattern(self, pat):
        k = pat[0]
        if k == "pid":
            return pat
        if k == "ptarget":
            return ("ptarget", self.assign_target(pat[1]))
        if k == "pobj":
            return ("pobj", [(key, self.pattern(t), self.expr(d) if d is not None else None) for key, t, d in pat[1]], pat[2])
        if k == "parr":
            return ("parr", [None if el is None else (self.pattern(el[0]), self.expr(el[1]) if el[1] is not None else None) for el in pat[1]],
                    self.pattern(pat[2]) if pat[2] is not None else None)
        if k in ("id",):
            return ("pid", pat[1])
        if k in ("mem", "idx"):
            return ("ptarget", self.assign_target(pat))
        raise SyntaxError(f"{self.path}: bad pattern {k}")

    # -- statements
    def block_needs_env(self, body):
        for st in body:
            k = st[0]
            if k in ("fdecl", "cdecl", "enum") or (k == "var" and st[1] != "var"):
                return True
            if k == "var":
                return True
        return False

    def block(self, body, new_env=True):
        hoisted = [(st[1], self.make_function(st[2])) for st in body if st[0] == "fdecl"]
        stmts = [self.stmt(st) for st in body if st[0] != "fdecl"]
        needs = new_env and (self.block_needs_env(body) or bool(hoisted))
        n = len(stmts)
        if not needs and not hoisted:
            if n == 1:
                return stmts[0]

            def run_plain(env):
                for s in stmts:
                    r = s(env)
                    if r is not None:
                        return r
                return None
            return run_plain

        def run(env):
            e = Env(env) if needs else env
            for name, mk in hoisted:
                e.vars[name] = mk(e)
            for s in stmts:
                r = s(e)
                if r is not None:
                    return r
            return None
        return run

    def stmt(self, st):
        k = st[0]
        if k == "expr":
            e = self.expr(st[1])

            def run_expr(env):
                e(env)
                return None
            return run_expr
        if k == "var":
            decls = []
            for pat, init in st[2]:
                decls.append((self.pattern(pat), self.expr(init) if init is not None else None))
            if len(decls) == 1 and decls[0][0][0] == "pid":
                name, init = decls[0][0][1], decls[0][1]
                if init is None:
                    def run_decl0(env):
                        env.vars[name] = UNDEF
                        return None
                    return run_decl0

                def run_decl1(env):
                    env.vars[name] = init(env)
                    return None
                return run_decl1

            def run_decl(env):
                for pat, init in decls:
                    bind_pattern(pat, init(env) if init is not None else UNDEF, env)
                return None
            return run_decl
        if k == "block":
            return self.block(st[1])
        if k == "if":
            c, a = self.expr(st[1]), self.stmt(st[2])
            b = self.stmt(st[3]) if st[3] is not None else None
            if b is None:
                def run_if(env):
                    if truthy(c(env)):
                        return a(env)
                    return None
                return run_if

            def run_ifelse(env):
                if truthy(c(env)):
                    return a(env)
                return b(env)
            return run_ifelse
        if k == "for":
            init = self.stmt(st[1]) if st[1] is not None else None
            test = self.expr(st[2]) if st[2] is not None else None
            upd = self.expr(st[3]) if st[3] is not None else None
            body = self.stmt(st[4])
            captures = st[1] is not None and st[1][0] == "var" and st[1][1] != "var" and self.contains_function(st[4])
            loop_vars = self.declared_names(st[1]) if captures else []

            def run_for(env):
                e = Env(env)
                if init is not None:
                    init(e)
                while True:
                    if test is not None and not truthy(test(e)):
                        break
                    r = body(e)
                    if r is not None:
                        if r is BREAK:
                            break
                        if r is not CONT:
                            return r
                    if captures:  # per-iteration binding for closures created in the body
                        e2 = Env(env)
                        for nm in loop_vars:
                            e2.vars[nm] = e.vars[nm]
                        e = e2
                    if upd is not None:
                        upd(e)
                return None
            return run_for
        if k == "forof" or k == "forin":
            pat, it, body = self.pattern(st[2]), self.expr(st[3]), self.stmt(st[4])
            is_in = k == "forin"

            def run_forof(env):
                seq = it(env)
                items = own_keys(seq) if is_in else iterate(seq)
                for x in items:
                    e = Env(env)
                    bind_pattern(pat, x, e)
                    r = body(e)
                    if r is not None:
                        if r is BREAK:
                            break
                        if r is not CONT:
                            return r
                return None
            return run_forof
        if k == "while":
            c, body = self.expr(st[1]), self.stmt(st[2])

            def run_while(env):
                while truthy(c(env)):
                    r = body(env)
                    if r is not None:
                        if r is BREAK:
                            break
                        if r is not CONT:
                            return r
                return None
            return run_while
        if k == "dowhile":
            body, c = self.stmt(st[1]), self.expr(st[2])

            def run_do(env):
                while True:
                    r = body(env)
                    if r is not None:
                        if r is BREAK:
                            break
                        if r is not CONT:
                            return r
                    if not truthy(c(env)):
                        break
                return None
            return run_do
        if k == "ret":
            if st[1] is None:
                return lambda env: ("r", UNDEF)
            e = self.expr(st[1])
            return lambda env: ("r", e(env))
        if k == "throw":
            e = self.expr(st[1])

            def run_throw(env):
                raise JSThrow(e(env))
            return run_throw
        if k == "break":
            return lambda env: BREAK
        if k == "cont":
            return lambda env: CONT
        if k == "empty":
            return lambda env: None
        if k == "switch":
            disc = self.expr(st[1])
            cases = [(self.expr(t) if t is not None else None, self.block(body, new_env=False)) for t, body in st[2]]
            default_at = next((i for i, (t, _) in enumerate(cases) if t is None), -1)

            def run_switch(env):
                v = disc(env)
                e = Env(env)
                start = -1
                for i, (t, _) in enumerate(cases):
                    if t is not None and strict_eq(v, t(e)):
                        start = i
                        break
                if start < 0:
                    start = default_at
                if start < 0:
                    return None
                for _, body in cases[start:]:
                    r = body(e)
                    if r is not None:
                        if r is BREAK:
                            return None
                        return r
                return None
            return run_switch
        if k == "try":
            blk = self.block(st[1])
            param = self.pattern(st[2]) if st[2] is not None else None
            handler = self.block(st[3]) if st[3] is not None else None
            fin = self.block(st[4]) if st[4] is not None else None

            def run_try(env):
                try:
                    try:
                        r = blk(env)
                    except JSThrow as ex:
                        if handler is None:
                            raise
                        e = Env(env)
                        if param is not None:
                            bind_pattern(param, ex.value, e)
                        r = handler(e)
                finally:
                    if fin is not None:
                        rf = fin(env)
                        if rf is not None:
                            return rf
                return r
            return run_try
        if k == "fdecl":
            mk = self.make_function(st[2])
            name = st[1]

            def run_fdecl(env):
                env.vars[name] = mk(env)
                return None
            return run_fdecl
        if k == "cdecl":
            mk = self.make_class(st[1], st[2], st[3])
            name = st[1]

            def run_cdecl(env):
                env.vars[name] = mk(env)
                return None
            return run_cdecl
        if k == "enum":
            name = st[1]
            members = [(key, self.expr(init) if init is not None else None) for key, init in st[2]]

            def run_enum(env):
                obj = {}
                e = Env(env)
                nxt = 0.0
                for key, init in members:
                    v = init(e) if init is not None else nxt
                    obj[key] = v
                    e.vars[key] = v
                    if type(v) is float:
                        obj[num_to_str(v)] = key
                        nxt = v + 1.0
                env.vars[name] = obj
                return None
            return run_enum
        if k in ("import", "export", "export_names", "export_star", "export_default", "export_default_decl"):
            raise SyntaxError(f"{self.path}: import / export below the top level")
        raise SyntaxError(f"{self.path}: unknown statement {k}")

    def contains_function(self, node):
        if isinstance(node, tuple):
            if node and node[0] in ("fn", "class", "fdecl", "cdecl"):
                return True
            return any(self.contains_function(x) for x in node)
        if isinstance(node, list):
            return any(self.contains_function(x) for x in node)
        return False

    # -- functions and classes
    def make_function(self, fn):
        _, params, body, is_arrow, is_expr, name, props = fn
        cparams = [(self.pattern(p), self.expr(d) if d is not None else None, rest) for p, d, rest in params]
        simple = None
        if all(p[0][0] == "pid" and p[1] is None and not p[2] for p in cparams):
            simple = [p[0][1] for p in cparams]
        cbody = self.expr(body) if is_expr else self.block(body, new_env=False)
        pprops = set(props)

        def mk(env):
            return JSFunction(cparams, cbody, env, is_arrow, is_expr, name, simple, pprops)
        return mk

    def make_class(self, name, members, parent=None):
        parent_expr = self.expr(parent) if parent is not None else None
        methods, statics_m, fields, statics_f, ctor = [], [], [], [], None
        for m in members:
            if m[0] == "method":
                _, key, fn, static = m
                if key == "constructor" and not static:
                    ctor = self.make_function(fn)
                else:
                    (statics_m if static else methods).append((key, self.make_function(fn)))
            else:
                _, key, init, static = m
                (statics_f if static else fields).append((key, self.expr(init) if init is not None else None))

        def mk(env):
            cls = JSClass(name or "")
            cenv = Env(env)
            if name:
                cenv.vars[name] = cls
            cenv.vars["%home"] = cls              # the class a method was defined in: what super.* is relative to
            if parent_expr is not None:
                cls.parent = parent_expr(env)
                if type(cls.parent) is not JSClass:
                    throw_type_error("Class extends value is not a class defined in the interpreted program")
            for key, f in methods:
                cls.methods[key] = f(cenv)
            if ctor is not None:
                cls.ctor = ctor(cenv)
            for key, init in fields:
                if init is None:
                    cls.fields.append((key, None))
                else:
                    def run_init(obj, init=init):
                        e = Env(cenv)
                        e.vars["this"] = obj
                        return init(e)
                    cls.fields.append((key, run_init))
            for key, f in statics_m:
                cls.statics[key] = BoundMethod(cls, f(cenv))
            for key, init in statics_f:
                e = Env(cenv)
                e.vars["this"] = cls
                cls.statics[key] = init(e) if init is not None else UNDEF
            return cls
        return mk

    # -- expressions
    def expr(self, e):
        k = e[0]
        m = getattr(self, "x_" + k, None)
        if m is None:
            raise SyntaxError(f"{self.path}: unknown expression node {k}")
        return m(e)

    def x_num(self, e):
        v = e[1]
        return lambda env: v

    x_str = x_num
    x_lit = x_num

    def x_tpl(self, e):
        parts = [self.expr(p) for p in e[1]]
        return lambda env: "".join(to_str(p(env)) for p in parts)

    def x_id(self, e):
        name = e[1]
        gvars = self.interp.globals.vars

        def lookup(env):
            s = env
            while s is not None:
                v = s.vars
                if name in v:
                    return v[name]
                s = s.parent
            raise JSThrow(JSObj(ERR.get("ReferenceError"), {"message": f"{name} is not defined", "name": "ReferenceError", "stack": ""}))
        return lookup

    def x_arr(self, e):
        elems = [None if el is None else (("s", self.expr(el[1])) if el[0] == "spread" else ("e", self.expr(el))) for el in e[1]]

        def build(env):
            out = []
            for el in elems:
                if el is None:
                    out.append(UNDEF)
                elif el[0] == "s":
                    out.extend(iterate(el[1](env)))
                else:
                    out.append(el[1](env))
            return out
        return build

    def x_obj(self, e):
        props = []
        for p in e[1]:
            if p[0] == "spread":
                props.append(("s", self.expr(p[1]), None))
            else:
                props.append(("p", self.expr(p[1]), self.expr(p[2])))

        def build(env):
            out = {}
            for kind, a, b in props:
                if kind == "s":
                    src = a(env)
                    if src is not None and src is not UNDEF:
                        for key in own_keys(src):
                            out[key] = get_prop(src, key)
                else:
                    out[prop_key(a(env))] = b(env)
            return out
        return build

    def x_fn(self, e):
        return self.make_function(e)

    def x_class(self, e):
        return self.make_class(e[1], e[2], e[3])

    def x_regex(self, e):
        source, flags = e[1], e[2]
        return lambda env: JSRegExp(source, flags)

    def x_super(self, e):
        raise SyntaxError(f"{self.path}: 'super' is only supported as super(...) and super.method(...)")

    def x_seq(self, e):
        parts = [self.expr(p) for p in e[1]]

        def run(env):
            v = UNDEF
            for p in parts:
                v = p(env)
            return v
        return run

    def x_cond(self, e):
        c, a, b = self.expr(e[1]), self.expr(e[2]), self.expr(e[3])
        return lambda env: a(env) if truthy(c(env)) else b(env)

    def x_log(self, e):
        op, a, b = e[1], self.expr(e[2]), self.expr(e[3])
        if op == "&&":
            def land(env):
                v = a(env)
                return b(env) if truthy(v) else v
            return land
        if op == "||":
            def lor(env):
                v = a(env)
                return v if truthy(v) else b(env)
            return lor

        def nullish(env):
            v = a(env)
            return b(env) if (v is None or v is UNDEF) else v
        return nullish

    def x_un(self, e):
        op = e[1]
        if op == "typeof":
            if e[2][0] == "id":
                inner = self.expr(e[2])

                def typeof_id(env):
                    try:
                        return js_typeof(inner(env))
                    except JSThrow:
                        return "undefined"
                return typeof_id
            a = self.expr(e[2])
            return lambda env: js_typeof(a(env))
        if op == "delete":
            t = e[2]
            if t[0] == "mem":
                o, name = self.expr(t[1]), t[2]
                return lambda env: _delete(o(env), name)
            if t[0] == "idx":
                o, i = self.expr(t[1]), self.expr(t[2])
                return lambda env: _delete(o(env), i(env))
            return lambda env: True
        a = self.expr(e[2])
        if op == "!":
            return lambda env: not truthy(a(env))
        if op == "-":
            def neg(env):
                v = a(env)
                return -v if type(v) is float else -to_number(v)
            return neg
        if op == "+":
            return lambda env: to_number(a(env))
        if op == "~":
            return lambda env: float(~to_int32(a(env)))
        if op == "void":
            def void(env):
                a(env)
                return UNDEF
            return void
        raise SyntaxError(f"{self.path}: unary {op}")

    def x_bin(self, e):
        op, a, b = e[1], self.expr(e[2]), self.expr(e[3])
        if op == "+":
            def add(env):
                x = a(env)
                y = b(env)
                if type(x) is float and type(y) is float:
                    return x + y
                return js_add(x, y)
            return add
        if op == "-":
            def sub(env):
                x = a(env)
                y = b(env)
                if type(x) is float and type(y) is float:
                    return x - y
                return to_number(x) - to_number(y)
            return sub
        if op == "*":
            def mul(env):
                x = a(env)
                y = b(env)
                if type(x) is float and type(y) is float:
                    return x * y
                return to_number(x) * to_number(y)
            return mul
        if op == "/":
            def div(env):
                x = a(env)
                y = b(env)
                if type(x) is not float:
                    x = to_number(x)
                if type(y) is not float:
                    y = to_number(y)
                if y != 0.0:
                    return x / y
                return js_div(x, y)
            return div
        if op == "%":
            return lambda env: js_mod(to_number(a(env)), to_number(b(env)))
        if op == "**":
            return lambda env: js_pow(to_number(a(env)), to_number(b(env)))
        if op in ("<", ">", "<=", ">="):
            import operator
            pyop = {"<": operator.lt, ">": operator.gt, "<=": operator.le, ">=": operator.ge}[op]

            def rel(env):
                x = a(env)
                y = b(env)
                if type(x) is float and type(y) is float:
                    return pyop(x, y)
                return js_compare(op, x, y)
            return rel
        if op == "===":
            def seq(env):
                x = a(env)
                y = b(env)
                if type(x) is float and type(y) is float:
                    return x == y
                return strict_eq(x, y)
            return seq
        if op == "!==":
            def sne(env):
                x = a(env)
                y = b(env)
                if type(x) is float and type(y) is float:
                    return x != y
                return not strict_eq(x, y)
            return sne
        if op == "==":
            return lambda env: loose_eq(a(env), b(env))
        if op == "!=":
            return lambda env: not loose_eq(a(env), b(env))
        if op == "&":
            return lambda env: float(to_int32(a(env)) & to_int32(b(env)))
        if op == "|":
            return lambda env: float(to_int32(a(env)) | to_int32(b(env)))
        if op == "^":
            return lambda env: float(to_int32(a(env)) ^ to_int32(b(env)))
        if op == "<<":
            return lambda env: float(to_int32(float((to_int32(a(env)) << (to_uint32(b(env)) & 31)) & 0xFFFFFFFF)))
        if op == ">>":
            return lambda env: float(to_int32(a(env)) >> (to_uint32(b(env)) & 31))
        if op == ">>>":
            return lambda env: float(to_uint32(a(env)) >> (to_uint32(b(env)) & 31))
        if op == "instanceof":
            return lambda env: _instanceof(a(env), b(env))
        if op == "in":
            return lambda env: _has_prop(b(env), a(env))
        raise SyntaxError(f"{self.path}: binary {op}")

    def x_mem(self, e):
        o, name, optional = self.expr(e[1]), e[2], e[3]
        if optional:
            def get_opt(env):
                v = o(env)
                return UNDEF if (v is None or v is UNDEF) else get_prop(v, name)
            return get_opt
        if name == "length":
            def get_len(env):
                v = o(env)
                t = type(v)
                if t is TypedArray:
                    return float(len(v.a))
                if t is list:
                    return float(len(v))
                return get_prop(v, name)
            return get_len

        def get(env):
            v = o(env)
            if type(v) is JSObj:
                p = v.props
                if name in p:
                    return p[name]
            return get_prop(v, name)
        return get

    def x_idx(self, e):
        o, i, optional = self.expr(e[1]), self.expr(e[2]), e[3]

        def get(env):
            v = o(env)
            if optional and (v is None or v is UNDEF):
                return UNDEF
            k = i(env)
            if type(k) is float:
                t = type(v)
                if t is TypedArray:
                    j = int(k)
                    if j == k and 0 <= j < len(v.a):
                        return float(v.a[j])
                    return UNDEF
                if t is list:
                    j = int(k)
                    if j == k and 0 <= j < len(v):
                        return v[j]
                    return UNDEF
            return get_prop(v, k)
        return get

    def assign_target(self, t):
        """-> setter(env, value)"""
        k = t[0]
        if k == "id":
            name = t[1]
            return lambda env, v: assign_var(env, name, v)
        if k == "mem":
            o, name = self.expr(t[1]), t[2]
            return lambda env, v: set_prop(o(env), name, v)
        if k == "idx":
            o, i = self.expr(t[1]), self.expr(t[2])
            return lambda env, v: set_prop(o(env), i(env), v)
        raise SyntaxError(f"{self.path}: bad assignment target {k}")

    def x_assign(self, e):
        op, target, rhs = e[1], e[2], self.expr(e[3])
        if target[0] in ("pobj", "parr"):
            pat = self.pattern(target)

            def destructure(env):
                v = rhs(env)
                bind_pattern(pat, v, env, declare=False)
                return v
            return destructure
        if op == "=":
            if target[0] == "id":
                name = target[1]

                def set_var(env):
                    v = rhs(env)
                    s = env
                    while s is not None:
                        d = s.vars
                        if name in d:
                            d[name] = v
                            return v
                        s = s.parent
                    assign_var(env, name, v)
                return set_var
            if target[0] == "mem":
                o, name = self.expr(target[1]), target[2]

                def set_mem(env):
                    obj = o(env)
                    v = rhs(env)
                    if type(obj) is JSObj:
                        obj.props[name] = v
                    else:
                        set_prop(obj, name, v)
                    return v
                return set_mem
            o, i = self.expr(target[1]), self.expr(target[2])

            def set_idx(env):
                obj = o(env)
                k = i(env)
                v = rhs(env)
                if type(obj) is TypedArray and type(k) is float:
                    j = int(k)
                    if j == k and 0 <= j < len(obj.a):
                        if type(v) is float and obj.a.typecode in "fd":
                            obj.a[j] = v
                        else:
                            obj.store(j, v)
                    return v
                set_prop(obj, k, v)
                return v
            return set_idx
        # compound assignment: evaluate the reference once
        binop = op[:-1]
        if binop in ("&&", "||", "??"):
            getter = self.expr(target)
            setter = self.assign_target(target)

            def logical_assign(env):
                cur = getter(env)
                if (binop == "&&" and truthy(cur)) or (binop == "||" and not truthy(cur)) or (binop == "??" and (cur is None or cur is UNDEF)):
                    v = rhs(env)
                    setter(env, v)
                    return v
                return cur
            return logical_assign
        combine = self.x_bin(("bin", binop, ("slot", 0), ("slot", 1)))
        if target[0] == "id":
            name = target[1]
            getter = self.x_id(target)

            def cvar(env):
                v = combine(_Slots(getter(env), rhs(env)))
                assign_var(env, name, v)
                return v
            return cvar
        if target[0] == "mem":
            o, name = self.expr(target[1]), target[2]

            def cmem(env):
                obj = o(env)
                v = combine(_Slots(get_prop(obj, name), rhs(env)))
                set_prop(obj, name, v)
                return v
            return cmem
        o, i = self.expr(target[1]), self.expr(target[2])

        def cidx(env):
            obj = o(env)
            k = i(env)
            v = combine(_Slots(get_prop(obj, k), rhs(env)))
            set_prop(obj, k, v)
            return v
        return cidx

    def x_slot(self, e):
        n = e[1]
        return (lambda s: s.a) if n == 0 else (lambda s: s.b)

    def x_upd(self, e):
        op, prefix, target = e[1], e[2], e[3]
        delta = 1.0 if op == "++" else -1.0
        if target[0] == "id":
            name = target[1]

            def upd_var(env):
                s = env
                while s is not None:
                    d = s.vars
                    if name in d:
                        old = d[name]
                        if type(old) is not float:
                            old = to_number(old)
                        new = old + delta
                        d[name] = new
                        return new if prefix else old
                    s = s.parent
                assign_var(env, name, UNDEF)
            return upd_var
        if target[0] == "mem":
            o, name = self.expr(target[1]), target[2]

            def upd_mem(env):
                obj = o(env)
                old = to_number(get_prop(obj, name))
                set_prop(obj, name, old + delta)
                return old + delta if prefix else old
            return upd_mem
        if target[0] == "idx":
            o, i = self.expr(target[1]), self.expr(target[2])

            def upd_idx(env):
                obj = o(env)
                k = i(env)
                old = to_number(get_prop(obj, k))
                set_prop(obj, k, old + delta)
                new = to_number(get_prop(obj, k)) if type(obj) is TypedArray else old + delta
                return new if prefix else old
            return upd_idx
        raise SyntaxError(f"{self.path}: bad update target")

    def args(self, args):
        if any(a[0] == "spread" for a in args):
            parts = [("s", self.expr(a[1])) if a[0] == "spread" else ("e", self.expr(a)) for a in args]

            def build(env):
                out = []
                for kind, f in parts:
                    if kind == "s":
                        out.extend(iterate(f(env)))
                    else:
                        out.append(f(env))
                return out
            return build
        fs = [self.expr(a) for a in args]
        n = len(fs)
        if n == 0:
            return lambda env: []
        if n == 1:
            f0 = fs[0]
            return lambda env: [f0(env)]
        if n == 2:
            f0, f1 = fs
            return lambda env: [f0(env), f1(env)]
        return lambda env: [f(env) for f in fs]

    def x_call(self, e):
        callee, args, optional = e[1], self.args(e[2]), e[3]
        if callee[0] == "super":
            look = self.x_id(("id", "%super_ctor"))
            return lambda env: look(env)(*args(env))
        if callee[0] == "mem" and callee[1][0] == "super":
            name = callee[2]
            home, this = self.x_id(("id", "%home")), self.x_id(("id", "this"))

            def call_super_method(env):
                parent = home(env).parent
                f = parent.find_method(name) if parent is not None else None
                if f is None:
                    throw_type_error(f"(intermediate value).{name} is not a function")
                return call_function(f, this(env), args(env))
            return call_super_method
        if callee[0] in ("mem", "idx"):
            o = self.expr(callee[1])
            key = (lambda env, name=callee[2]: name) if callee[0] == "mem" else self.expr(callee[2])
            opt_member = callee[3]

            def call_method(env):
                obj = o(env)
                if opt_member and (obj is None or obj is UNDEF):
                    return UNDEF
                k = key(env)
                t = type(obj)
                if t is JSObj:
                    f = obj.props.get(k)
                    if f is None:
                        f = obj.cls.find_method(k) if (obj.cls is not None and type(k) is str) else None
                        if f is None:
                            f = get_prop(obj, k)
                elif t is Native:
                    f = obj.props.get(k, UNDEF)
                    if f is UNDEF or f is None:
                        if optional:
                            return UNDEF
                        throw_type_error(f"{obj.name}.{to_str(k)} is not a function")
                    a = args(env)
                    return f(*a) if not isinstance(f, (JSFunction, BoundMethod, Native, JSClass)) else call_function(f, obj, a)
                else:
                    f = get_prop(obj, k)
                if f is UNDEF or f is None:
                    if optional:
                        return UNDEF
                    throw_type_error(f"{to_str(k)} is not a function")
                a = args(env)
                if type(f) is JSFunction:
                    return call_function(f, obj, a)
                if type(f) is BoundMethod:
                    fn = f.fn
                    if type(fn) is JSFunction:
                        return call_function(fn, f.this, a)
                    return fn(f.this, *a)
                return call_function(f, obj, a)
            return call_method
        f = self.expr(callee)

        def call(env):
            fn = f(env)
            if optional and (fn is None or fn is UNDEF):
                return UNDEF
            return call_function(fn, UNDEF, args(env))
        return call

    def x_new(self, e):
        callee, args = self.expr(e[1]), self.args(e[2])
        return lambda env: construct(callee(env), args(env))


class _Slots:
    __slots__ = ("a", "b")

    def __init__(self, a, b):
        self.a, self.b = a, b


def _delete(obj, key):
    if type(obj) is dict:
        obj.pop(prop_key(key), None)
    elif type(obj) is JSObj:
        obj.props.pop(prop_key(key), None)
    elif type(obj) is list:
        i = idx_int(key)
        if 0 <= i < len(obj):
            obj[i] = UNDEF
    return True


def _instanceof(v, cls):
    if type(cls) is JSClass:
        if type(v) is JSObj:
            c = v.cls
            while c is not None:
                if c is cls:
                    return True
                c = c.parent
            return cls.name == "Error" and v.cls is not None and v.cls.name.endswith("Error")
        return False
    if type(cls) is Native:
        if cls.name == "Array":
            return type(v) is list
        if cls.name in KIND_CODE:
            return type(v) is TypedArray and v.kind == cls.name
        if cls.name == "Map":
            return type(v) is JSMap
        if cls.name == "Set":
            return type(v) is JSSet
        if cls.name == "Object":
            return isinstance(v, (dict, JSObj, list, TypedArray, JSMap, JSSet))
    return False


def _has_prop(obj, key):
    if type(obj) is dict:
        return prop_key(key) in obj
    if type(obj) is JSObj:
        k = prop_key(key)
        return k in obj.props or (obj.cls is not None and obj.cls.find_method(k) is not None)
    if type(obj) in (list, TypedArray):
        n = len(obj.a) if type(obj) is TypedArray else len(obj)
        return 0 <= idx_int(key) < n or key == "length"
    throw_type_error("Cannot use 'in' operator on a primitive")
