"""hs_eval -- a small evaluator for the subset of Haskell the reference's Level-1 path is written in.
TEST INFRASTRUCTURE (only tests/ and tests/golden/make_ref_golden.py use it; the product never does).

Why it exists.  The reference (dlewissandy/lambda-blas) is Haskell, and neither GHC nor stack/cabal exist in
this image or on the GPU box, so the reference cannot be compiled or run; it also holds no golden vectors
(its tests are unseeded QuickCheck properties against Fortran netlib, which is absent too).  What CAN be done
here is to evaluate the reference's own SOURCE TEXT: this module lexes (with the layout rule), parses and
evaluates

    /root/reference/src/Numerical/BLAS/Types.hs   (the ADTs)
    /root/reference/src/Numerical/BLAS/Util.hs    (sampleElems, sampleIndices, updateElems, ...)
    /root/reference/src/Numerical/BLAS/Single.hs  (dot, asum, nrm2, iamax, sdsdot, scal, copy, swap, axpy,
                                                   rot, rotm, rotg, rotmg, checkscale)

as they lie there -- nothing of those files is copied into this repository -- and produces the values the
reference's functions return.  tests/golden/make_ref_golden.py runs it in the build container and commits the
input/output vectors (tests/golden/ref_*.json); oracle/lblas_oracle.c and the CUDA path are then checked against
those vectors bit for bit (tests/test_ref_golden.py, tests/test_gpu_ref_golden.py).

What is NOT the reference here, and is restated from the documentation of the pinned dependencies
(stack.yaml:18 `resolver: lts-7.14` = GHC 8.0.1 / base-4.9.0.0 / vector-0.11.0.0; none of them vendored):
  * the primitives of `base`: IEEE binary32 / binary64 arithmetic, one correctly rounded operation each
    (numpy scalars), `compare` (LT if x<y, EQ if x==y, else GT -- so a NaN operand gives GT), `abs`
    (0 for x==0, x for x>0, else negate x), `signum`, `(^)` by repeated squaring (GHC.Real), literals by
    `fromInteger` / `fromRational` (correctly rounded), `float2Double#` / `double2Float#`;
  * the combinators of `Data.Vector.Storable` the reference calls (foldl', zipWith, map, unsafeTake,
    unsafeIndex, replicate, reverse, unsafeUpdate_, maxIndexBy) and the stream constructors it uses to build
    index vectors (Stream/Step/Exact, fromStream, unstream).  Vectors are evaluated as FUSED pipelines, the way
    -O2 (lambda-blas.cabal:48) compiles them: `unstream` is lazy, zips stop at the shorter side, unsafeUpdate_
    clones the destination and applies the writes in stream order, maxIndexBy keeps the earlier element unless
    `cmp earlier later == LT`.
Evaluation is call-by-need (thunks for arguments and let/where bindings), bang patterns force.

The evaluator covers what those three files use: module/import headers (skipped), data declarations with
records, type signatures (skipped), function clauses with guards and where, let, case, if, lambdas, operator
sections, tuples, as-/bang-/view-/literal-/constructor patterns, RecordWildCards construction, MagicHash names.
It is not a general Haskell implementation; anything outside the subset raises HsError.
"""
from __future__ import annotations

import re
import sys
import threading
from fractions import Fraction

import numpy as np


class HsError(Exception):
    pass


# --------------------------------------------------------------------------------------------------
# lexer
# --------------------------------------------------------------------------------------------------
KEYWORDS = {"module", "where", "import", "qualified", "as", "let", "in", "case", "of", "if", "then", "else",
            "data", "deriving", "instance", "type", "newtype", "class", "do", "hiding", "infixl", "infixr", "infix"}
SYMCHARS = set("!#$%&*+./<=>?@\\^|-~:")
RESERVED_OPS = {"..", ":", "::", "=", "\\", "|", "<-", "->", "@", "~", "=>"}

_re_qual = re.compile(r"(?:[A-Z][A-Za-z0-9_']*\.)+(?=[A-Za-z_])")
_re_var = re.compile(r"[a-z_][A-Za-z0-9_']*#*")
_re_con = re.compile(r"[A-Z][A-Za-z0-9_']*#*")
_re_num = re.compile(r"\d+(?:\.\d+)?(?:[eE][-+]?\d+)?")


class Tok:
    __slots__ = ("kind", "val", "line", "col", "first", "extra")

    def __init__(self, kind, val, line, col, first=False, extra=None):
        self.kind, self.val, self.line, self.col, self.first, self.extra = kind, val, line, col, first, extra

    def __repr__(self):
        return f"{self.kind}:{self.val}@{self.line}:{self.col}"


def lex(src: str):
    toks = []
    i, n, line, col = 0, len(src), 1, 1
    line_has_tok = False

    def adv(k):
        nonlocal i, line, col, line_has_tok
        for _ in range(k):
            if src[i] == "\n":
                line += 1
                col = 1
                line_has_tok = False
            elif src[i] == "\t":
                col += 8 - (col - 1) % 8
            else:
                col += 1
            i += 1

    def emit(kind, val, length, extra=None):
        nonlocal line_has_tok
        toks.append(Tok(kind, val, line, col, not line_has_tok, extra))
        line_has_tok = True
        adv(length)

    while i < n:
        c = src[i]
        if c in " \t\r\n":
            adv(1)
            continue
        if src.startswith("{-", i):                     # block comment / pragma, nested
            depth, j = 1, i + 2
            while j < n and depth:
                if src.startswith("{-", j):
                    depth += 1
                    j += 2
                elif src.startswith("-}", j):
                    depth -= 1
                    j += 2
                else:
                    j += 1
            adv(j - i)
            continue
        if c == "-" and src.startswith("--", i):
            j = i
            while j < n and src[j] == "-":
                j += 1
            if j >= n or src[j] not in SYMCHARS:        # a run of dashes not followed by a symbol: line comment
                while i < n and src[i] != "\n":
                    adv(1)
                continue
        m = _re_qual.match(src, i)
        qual = m.group(0) if m else ""
        j = i + len(qual)
        m = _re_var.match(src, j)
        if m:
            name = m.group(0)
            if not qual and name in KEYWORDS:
                emit("kw", name, len(name))
            else:
                emit("varid", qual + name, len(qual) + len(name))
            continue
        m = _re_con.match(src, j)
        if m:
            emit("conid", qual + m.group(0), len(qual) + len(m.group(0)))
            continue
        m = _re_num.match(src, i)
        if m and c.isdigit():
            txt = m.group(0)
            isfrac = any(ch in txt for ch in ".eE")
            emit("num", Fraction(txt), len(txt), isfrac)
            continue
        if c in "()[],;{}`":
            emit("special", c, 1)
            continue
        if c in SYMCHARS:
            j = i
            while j < n and src[j] in SYMCHARS:
                j += 1
            emit("op", src[i:j], j - i)
            continue
        if c == "'":                                     # a character literal (an identifier's primes were consumed above)
            m = re.compile(r"'(?:\\[^']+|\\'|[^'\\])'").match(src, i)
            if not m:
                raise HsError(f"lex: stray quote at {line}:{col}")
            emit("char", m.group(0)[1:-1], len(m.group(0)))
            continue
        if c == '"':
            j = i + 1
            while src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            emit("str", src[i + 1:j], j + 1 - i)
            continue
        raise HsError(f"lex: unexpected {c!r} at {line}:{col}")
    return toks


# --------------------------------------------------------------------------------------------------
# layout rule (Haskell 2010 report 10.3, with the parse-error(t) rule implemented for `in`, closing
# brackets and commas -- the only places the three files need it)
# --------------------------------------------------------------------------------------------------
def layout(toks):
    out = []
    ctx = []            # implicit blocks: dict(col, depth, kind); explicit braces only count as bracket depth
    pend = []           # `if` / `case` keywords whose `else` / `of` has not been seen: [keyword, len(ctx) then, bracket depth]
    depth = 0
    expect = None

    def close():
        b = ctx.pop()
        out.append(Tok("special", "}", out[-1].line if out else 0, 0, extra="virtual"))
        for q in pend:
            q[1] = min(q[1], len(ctx))
        return b

    for t in toks:
        closed_let = False
        if expect is not None:
            kind, expect = expect, None
            if t.kind == "special" and t.val == "{":
                raise HsError("explicit layout braces are outside the subset")
            out.append(Tok("special", "{", t.line, t.col, extra="virtual"))
            if ctx and t.first and t.col <= ctx[-1]["col"]:          # an empty block: `where` followed by a less indented line
                out.append(Tok("special", "}", t.line, t.col, extra="virtual"))
            else:
                ctx.append({"col": t.col, "depth": depth, "kind": kind})
                t = Tok(t.kind, t.val, t.line, t.col, False, t.extra)      # the block's first token: no `;` before it
        if t.first:
            while ctx and (t.col < ctx[-1]["col"] or (t.kind == "kw" and t.val == "where" and ctx[-1]["kind"] in ("of", "do")
                                                      and t.col == ctx[-1]["col"])):      # parse-error(t): no alternative / statement starts with `where`
                closed_let = closed_let or close()["kind"] == "let"
            if ctx and t.col == ctx[-1]["col"]:
                out.append(Tok("special", ";", t.line, t.col, extra="virtual"))
        if t.kind == "kw" and t.val == "in" and not closed_let:
            # parse-error(t): ends the innermost `let` block that is still open inside the current brackets (a `let` whose
            # block the indentation of this very token has just closed needs nothing more; `let` statements of a
            # do block never meet an `in`)
            k = len(ctx) - 1
            while k >= 0 and ctx[k]["depth"] == depth and ctx[k]["kind"] != "let":
                k -= 1
            if k >= 0 and ctx[k]["depth"] == depth and ctx[k]["kind"] == "let":
                while len(ctx) > k:
                    close()
        if t.kind == "kw" and t.val in ("if", "case"):
            pend.append([t.val, len(ctx), depth])
        elif t.kind == "kw" and t.val in ("then", "else", "of"):
            # parse-error(t): implicit blocks opened since the matching `if` / `case` (a one-line `do` inside a branch, a
            # `case` inside a scrutinee) end here
            want = "case" if t.val == "of" else "if"
            while pend and (pend[-1][0] != want or pend[-1][2] > depth):
                pend.pop()
            if pend:
                while len(ctx) > pend[-1][1] and ctx[-1]["depth"] == depth:
                    close()
                if t.val != "then":
                    pend.pop()
        if t.kind == "special" and t.val in ")]":
            while ctx and ctx[-1]["depth"] == depth:
                close()
            depth -= 1
        elif t.kind == "special" and t.val == "}":
            while ctx and ctx[-1]["depth"] == depth:
                close()
            depth -= 1
        elif t.kind == "special" and t.val == "," and depth > 0:
            while ctx and ctx[-1]["depth"] == depth:
                close()
        out.append(t)
        if t.kind == "special" and t.val in "([{":
            depth += 1
        if t.kind == "kw" and t.val in ("where", "let", "of", "do"):
            expect = t.val
    while ctx:
        close()
    return out


# --------------------------------------------------------------------------------------------------
# parser
# --------------------------------------------------------------------------------------------------
FIXITY = {"$": (0, "r"), "||": (2, "r"), "&&": (3, "r"),
          "==": (4, "n"), "/=": (4, "n"), "<": (4, "n"), "<=": (4, "n"), ">": (4, "n"), ">=": (4, "n"),
          "+": (6, "l"), "-": (6, "l"), "*": (7, "l"), "/": (7, "l"), "^": (8, "r"), ".": (9, "r")}


class Parser:
    def __init__(self, toks, fname="?"):
        self.t = toks
        self.i = 0
        self.fname = fname

    # -- token helpers
    def peek(self, k=0):
        j = self.i + k
        return self.t[j] if j < len(self.t) else Tok("eof", None, -1, -1)

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def is_(self, kind, val=None, k=0):
        t = self.peek(k)
        return t.kind == kind and (val is None or t.val == val)

    def accept(self, kind, val=None):
        if self.is_(kind, val):
            return self.next()
        return None

    def expect(self, kind, val=None):
        t = self.next()
        if t.kind != kind or (val is not None and t.val != val):
            raise HsError(f"{self.fname}: expected {kind} {val!r}, got {t!r}")
        return t

    def skip_decl(self):
        """skip to the `;` / `}` that ends the current declaration"""
        d = 0
        while True:
            t = self.peek()
            if t.kind == "eof":
                return
            if t.kind == "special":
                if t.val in "([{":
                    d += 1
                elif t.val in ")]}":
                    if d == 0:
                        return
                    d -= 1
                elif t.val == ";" and d == 0:
                    return
            self.next()

    # -- module
    def parse_module(self):
        decls, datas = [], {}
        if self.is_("kw", "module"):
            while not self.is_("kw", "where"):
                self.next()
            self.next()
        else:
            raise HsError("module header expected")
        self.expect("special", "{")
        while not self.is_("special", "}") and not self.is_("eof"):
            if self.accept("special", ";"):
                continue
            if self.is_("kw", "import") or self.is_("kw", "instance") or self.is_("kw", "type"):
                self.skip_decl()
            elif self.is_("kw", "data"):
                self.parse_data(datas)
            else:
                d = self.parse_decl()
                if d is not None:
                    decls.append(d)
        return decls, datas

    def parse_data(self, datas):
        self.expect("kw", "data")
        while not self.is_("op", "="):
            self.next()
        self.next()
        while True:
            con = self.expect("conid").val
            fields = None
            arity = 0
            if self.is_("special", "{"):
                self.next()
                fields = []
                while not self.is_("special", "}"):
                    fields.append(self.expect("varid").val)
                    if self.accept("special", ","):
                        continue
                    self.expect("op", "::")
                    d = 0
                    while True:
                        t = self.peek()
                        if t.kind == "special" and t.val in "([":
                            d += 1
                        elif t.kind == "special" and t.val in ")]":
                            d -= 1
                        elif d == 0 and t.kind == "special" and t.val in ",}":
                            break
                        self.next()
                    self.accept("special", ",")
                self.next()
                arity = len(fields)
            else:
                d = 0
                while True:
                    t = self.peek()
                    if d == 0 and (t.kind == "eof" or (t.kind == "op" and t.val == "|") or (t.kind == "kw" and t.val == "deriving")
                                   or (t.kind == "special" and t.val in ";}")):
                        break
                    if t.kind == "special" and t.val in "([":
                        if d == 0:
                            arity += 1
                        d += 1
                    elif t.kind == "special" and t.val in ")]":
                        d -= 1
                    elif d == 0 and t.kind in ("varid", "conid"):
                        arity += 1
                    self.next()
            datas[con] = {"fields": fields, "arity": arity}
            if not self.accept("op", "|"):
                break
        self.skip_decl()

    # -- declarations
    def parse_block(self):
        self.expect("special", "{")
        decls = []
        while not self.is_("special", "}"):
            if self.accept("special", ";"):
                continue
            d = self.parse_decl()
            if d is not None:
                decls.append(d)
        self.next()
        return decls

    def parse_decl(self):
        # type signature:  v1 [, v2 ...] :: type
        k = 0
        if self.is_("varid"):
            k = 1
            while self.is_("special", ",", k) and self.is_("varid", None, k + 1):
                k += 2
            if self.is_("op", "::", k):
                self.skip_decl()
                return None
        if self.is_("kw", "infixl") or self.is_("kw", "infixr") or self.is_("kw", "infix"):
            self.skip_decl()
            return None
        if self.is_("varid") and not self.is_("op", "@", 1):
            name = self.next().val
            pats = []
            while not (self.is_("op", "=") or self.is_("op", "|")):
                pats.append(self.parse_apat())
            rhs, wh = self.parse_rhs("=")
            return ("fun", name, pats, rhs, wh)
        pat = self.parse_pat()
        rhs, wh = self.parse_rhs("=")
        return ("patbind", pat, rhs, wh)

    def parse_rhs(self, eq):
        """guarded right-hand sides + optional where: ([(guard|None, exp)], decls)"""
        rhs = []
        if self.is_("op", "|"):
            while self.accept("op", "|"):
                g = self.parse_exp()
                self.expect("op", eq)
                rhs.append((g, self.parse_exp()))
        else:
            self.expect("op", eq)
            rhs.append((None, self.parse_exp()))
        wh = []
        if self.accept("kw", "where"):
            wh = self.parse_block()
        return rhs, wh

    # -- patterns
    def parse_pat(self):
        if self.is_("conid"):
            con = self.next().val
            args = []
            while self.apat_start():
                args.append(self.parse_apat())
            return ("pcon", con, args)
        if self.is_("op", "-") and self.is_("num", None, 1):
            self.next()
            return ("plit", -self.next().val)
        return self.parse_apat()

    def apat_start(self):
        t = self.peek()
        return t.kind in ("varid", "conid", "num") or (t.kind == "special" and t.val == "(") or (t.kind == "op" and t.val in ("!", "~"))

    def parse_apat(self):
        t = self.next()
        if t.kind == "op" and t.val == "!":
            return ("pbang", self.parse_apat())
        if t.kind == "varid":
            if t.val == "_":
                return ("pwild",)
            if self.accept("op", "@"):
                return ("pas", t.val, self.parse_apat())
            return ("pvar", t.val)
        if t.kind == "num":
            return ("plit", t.val)
        if t.kind == "conid":
            return ("pcon", t.val, [])
        if t.kind == "special" and t.val == "(":
            if self.accept("special", ")"):
                return ("ptuple", [])
            # view pattern?  look for `->` at depth 0 before the matching `)`
            d, k = 0, 0
            view = False
            while True:
                u = self.peek(k)
                if u.kind == "eof":
                    break
                if u.kind == "special" and u.val in "([{":
                    d += 1
                elif u.kind == "special" and u.val in ")]}":
                    if d == 0:
                        break
                    d -= 1
                elif d == 0 and u.kind == "op" and u.val == "->":
                    view = True
                    break
                elif d == 0 and u.kind == "special" and u.val == ",":
                    break
                k += 1
            if view:
                e = self.parse_exp()
                self.expect("op", "->")
                p = self.parse_pat()
                self.expect("special", ")")
                return ("pview", e, p)
            ps = [self.parse_pat()]
            while self.accept("special", ","):
                ps.append(self.parse_pat())
            self.expect("special", ")")
            return ps[0] if len(ps) == 1 else ("ptuple", ps)
        raise HsError(f"{self.fname}: pattern expected at {t!r}")

    # -- expressions
    def parse_exp(self):
        return self.parse_opexp(0)

    def cur_op(self):
        """the binary operator at the cursor: (name, prec, assoc, ntokens) or None"""
        t = self.peek()
        if t.kind == "op" and t.val in FIXITY:
            p, a = FIXITY[t.val]
            return t.val, p, a, 1
        if t.kind == "special" and t.val == "`" and self.is_("special", "`", 2):
            return self.peek(1).val, 9, "l", 3
        return None

    def parse_opexp(self, minp):
        if self.is_("op", "-") and minp <= 6:
            self.next()
            left = ("neg", self.parse_opexp(7))
        else:
            left = self.parse_exp10()
        while True:
            o = self.cur_op()
            if o is None:
                return left
            name, p, a, nt = o
            if p < minp:
                return left
            # a closing section `(e op)` : operator directly followed by `)`
            if self.is_("special", ")", nt):
                return left
            self.i += nt
            right = self.parse_opexp(p + 1 if a in "ln" else p)
            left = ("binop", name, left, right)

    def parse_exp10(self):
        t = self.peek()
        if t.kind == "op" and t.val == "\\":
            self.next()
            pats = []
            while not self.is_("op", "->"):
                pats.append(self.parse_apat())
            self.next()
            return ("lam", pats, self.parse_exp())
        if t.kind == "kw":
            if t.val == "let":
                self.next()
                decls = self.parse_block()
                self.expect("kw", "in")
                return ("let", decls, self.parse_exp())
            if t.val == "if":
                self.next()
                c = self.parse_exp()
                self.accept("special", ";")
                self.expect("kw", "then")
                a = self.parse_exp()
                self.accept("special", ";")
                self.expect("kw", "else")
                return ("if", c, a, self.parse_exp())
            if t.val == "case":
                self.next()
                scrut = self.parse_exp()
                self.expect("kw", "of")
                self.expect("special", "{")
                alts = []
                while not self.is_("special", "}"):
                    if self.accept("special", ";"):
                        continue
                    p = self.parse_pat()
                    rhs, wh = self.parse_rhs("->")
                    alts.append((p, rhs, wh))
                self.next()
                return ("case", scrut, alts)
        f = self.parse_aexp()
        while self.aexp_start():
            f = ("app", f, self.parse_aexp())
        return f

    def aexp_start(self):
        t = self.peek()
        return t.kind in ("varid", "conid", "num", "str") or (t.kind == "special" and t.val in "([")

    def parse_aexp(self):
        t = self.next()
        if t.kind == "varid":
            return ("var", t.val)
        if t.kind == "num":
            return ("lit", t.val, t.extra)
        if t.kind == "conid":
            if self.is_("special", "{") and self.peek().extra != "virtual":
                self.next()
                if self.accept("op", ".."):
                    self.expect("special", "}")
                    return ("reccon", t.val, None)
                fs = []
                while not self.is_("special", "}"):
                    fn = self.expect("varid").val
                    self.expect("op", "=")
                    fs.append((fn, self.parse_exp()))
                    self.accept("special", ",")
                self.next()
                return ("reccon", t.val, fs)
            return ("con", t.val)
        if t.kind == "special" and t.val == "(":
            if self.accept("special", ")"):
                return ("tuple", [])
            # (op) and right sections (op e)
            o = self.cur_op()
            if o is not None and not (o[0] == "-" and not self.is_("special", ")", 1)):
                name, p, a, nt = o
                self.i += nt
                if self.accept("special", ")"):
                    return ("var", "(" + name + ")")
                e = self.parse_opexp(p + 1 if a in "ln" else p)
                self.expect("special", ")")
                return ("rsec", name, e)
            e = self.parse_exp()
            o = self.cur_op()
            if o is not None and self.is_("special", ")", o[3]):      # left section (e op)
                self.i += o[3]
                self.expect("special", ")")
                return ("lsec", o[0], e)
            if self.accept("op", "::"):
                ty = []
                d = 0
                while not (d == 0 and self.is_("special", ")")):
                    u = self.next()
                    if u.kind == "special" and u.val == "(":
                        d += 1
                    if u.kind == "special" and u.val == ")":
                        d -= 1
                    ty.append(u.val)
                self.next()
                return ("typed", e, " ".join(str(x) for x in ty))
            if self.is_("special", ","):
                es = [e]
                while self.accept("special", ","):
                    es.append(self.parse_exp())
                self.expect("special", ")")
                return ("tuple", es)
            self.expect("special", ")")
            return e
        raise HsError(f"{self.fname}: expression expected at {t!r}")


# --------------------------------------------------------------------------------------------------
# values
# --------------------------------------------------------------------------------------------------
F32, F64 = np.float32, np.float64


class Lit:
    """A numeric literal whose type is not fixed yet (`fromInteger` / `fromRational` not applied): exact."""
    __slots__ = ("q", "frac")

    def __init__(self, q, frac=False):
        self.q, self.frac = Fraction(q), frac

    def __repr__(self):
        return f"Lit({self.q})"


def _round_to(q: Fraction, ty):
    """fromRational: the value of type `ty` nearest to q, ties to even"""
    d = float(q)                      # correctly rounded to binary64 (CPython int/int true division)
    if ty is F64:
        return F64(d)
    f = F32(d)
    if not np.isfinite(f):
        return f
    # double rounding can be off by one ulp: pick the nearest of the three neighbours exactly
    cands = {float(f), float(np.nextafter(f, F32(np.inf))), float(np.nextafter(f, F32(-np.inf)))}
    best = None
    for c in cands:
        if not np.isfinite(c):
            continue
        err = abs(Fraction(c) - q)
        even = (np.float32(c).view(np.uint32) & 1) == 0
        if best is None or err < best[0] or (err == best[0] and even):
            best = (err, even, c)
    return F32(best[2])


def to_type(v, ty):
    """give a literal the type `ty` (int, F32 or F64)"""
    if isinstance(v, Lit):
        if ty is int:
            if v.frac or v.q.denominator != 1:
                raise HsError(f"fractional literal {v.q} used as an Int")
            return int(v.q)
        return _round_to(v.q, ty)
    return v


def num_type(v):
    if isinstance(v, Lit):
        return None
    if isinstance(v, bool):
        raise HsError("Bool used as a number")
    if isinstance(v, int):
        return int
    if isinstance(v, F32):
        return F32
    if isinstance(v, F64):
        return F64
    raise HsError(f"not a number: {v!r}")


def unify(a, b):
    ta, tb = num_type(a), num_type(b)
    if ta is None and tb is None:
        return a, b, None
    if ta is None:
        return to_type(a, tb), b, tb
    if tb is None:
        return a, to_type(b, ta), ta
    if ta is not tb:
        raise HsError(f"type mismatch: {ta} vs {tb}")
    return a, b, ta


class Con:
    __slots__ = ("name", "args")

    def __init__(self, name, args=()):
        self.name, self.args = name, list(args)

    def __repr__(self):
        return f"{self.name}{self.args}"


class Thunk:
    __slots__ = ("expr", "env", "val", "state")

    def __init__(self, expr, env):
        self.expr, self.env, self.val, self.state = expr, env, None, 0


class Closure:
    __slots__ = ("clauses", "arity", "env", "args", "name")

    def __init__(self, clauses, arity, env, args=(), name="\\"):
        self.clauses, self.arity, self.env, self.args, self.name = clauses, arity, env, list(args), name


class Builtin:
    __slots__ = ("fn", "arity", "args", "name")

    def __init__(self, name, arity, fn, args=()):
        self.fn, self.arity, self.args, self.name = fn, arity, list(args), name


class Vec:
    """A storable vector: a materialised list of forced elements, or a fused pipeline (a generator factory)."""
    __slots__ = ("items", "gen")
    LIMIT = 10_000_000

    def __init__(self, items=None, gen=None):
        self.items, self.gen = items, gen

    def iter(self):
        return iter(self.items) if self.items is not None else self.gen()

    def force(self):
        if self.items is None:
            out = []
            for x in self.gen():
                out.append(x)
                if len(out) > Vec.LIMIT:
                    raise HsError("vector pipeline does not terminate")
            self.items, self.gen = out, None
        return self.items


class Env:
    __slots__ = ("d", "parent")

    def __init__(self, parent=None):
        self.d, self.parent = {}, parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.d:
                return e.d[name]
            e = e.parent
        raise HsError(f"unbound name {name}")


# --------------------------------------------------------------------------------------------------
# evaluator
# --------------------------------------------------------------------------------------------------
class MatchFail(Exception):
    pass


class Interp:
    def __init__(self):
        self.glob = Env()
        self.datas = {}
        self._install_prelude()

    # ---- loading
    def load(self, path):
        with open(path) as f:
            self.load_source(f.read(), path)

    def load_source(self, src, path="<source>"):
        decls, datas = Parser(layout(lex(src)), path).parse_module()
        self.datas.update(datas)
        self.bind_decls(decls, self.glob)

    def bind_decls(self, decls, env):
        funs = {}
        for d in decls:
            if d[0] == "fun":
                _, name, pats, rhs, wh = d
                funs.setdefault(name, []).append((pats, rhs, wh))
            else:
                _, pat, rhs, wh = d
                th = Thunk(("rhs", rhs, wh), env)
                self.bind_lazy_pattern(pat, th, env)
        for name, clauses in funs.items():
            ar = len(clauses[0][0])
            if any(len(c[0]) != ar for c in clauses):
                raise HsError(f"clauses of {name} differ in arity")
            if ar == 0:
                env.d[name] = Thunk(("rhs", clauses[0][1], clauses[0][2]), env)
            else:
                env.d[name] = Closure(clauses, ar, env, name=name)

    def bind_lazy_pattern(self, pat, th, env):
        """pattern bindings are lazy: every variable of the pattern becomes a thunk that matches on demand"""
        names = []

        def collect(p):
            k = p[0]
            if k == "pvar":
                names.append(p[1])
            elif k == "pas":
                names.append(p[1])
                collect(p[2])
            elif k == "pbang":
                collect(p[1])
            elif k in ("ptuple",):
                for q in p[1]:
                    collect(q)
            elif k == "pcon":
                for q in p[2]:
                    collect(q)
            elif k == "pview":
                collect(p[2])
        collect(pat)
        for nm in names:
            env.d[nm] = Thunk(("patsel", pat, th, nm), env)

    # ---- forcing
    def force(self, v):
        while isinstance(v, Thunk):
            if v.state == 2:
                v = v.val
                continue
            if v.state == 1:
                raise HsError("<<loop>>")
            v.state = 1
            r = self.eval(v.expr, v.env)
            while isinstance(r, Thunk):
                r = self.force(r)
            v.val, v.state, v.expr, v.env = r, 2, None, None
            v = r
        return v

    def delay(self, e, env):
        k = e[0]
        if k == "lit":
            return Lit(e[1], e[2])
        if k == "var":
            try:
                return env.lookup(e[1])
            except HsError:
                return Thunk(e, env)
        return Thunk(e, env)

    # ---- application
    def apply(self, f, arg):
        f = self.force(f)
        if isinstance(f, Closure):
            args = f.args + [arg]
            if len(args) < f.arity:
                return Closure(f.clauses, f.arity, f.env, args, f.name)
            return self.call_closure(f, args)
        if isinstance(f, Builtin):
            args = f.args + [arg]
            if len(args) < f.arity:
                return Builtin(f.name, f.arity, f.fn, args)
            return f.fn(*args)
        if isinstance(f, Con):
            return Con(f.name, f.args + [arg])
        raise HsError(f"cannot apply {f!r}")

    def apply_n(self, f, *args):
        for a in args:
            f = self.apply(f, a)
        return f

    def call_closure(self, f, args):
        for pats, rhs, wh in f.clauses:
            env = Env(f.env)
            try:
                for p, a in zip(pats, args):
                    self.match(p, a, env)
            except MatchFail:
                continue
            r = self.eval_rhs(rhs, wh, env, fall=True)
            if r is not MatchFail:
                return r
        raise HsError(f"non-exhaustive patterns in {f.name}")

    def eval_rhs(self, rhs, wh, env, fall=False):
        if wh:
            env = Env(env)
            self.bind_decls(wh, env)
        for g, e in rhs:
            if g is None or self.truth(self.force(self.eval(g, env))):
                return self.eval(e, env)
        if fall:
            return MatchFail
        raise HsError("no guard matched")

    @staticmethod
    def truth(v):
        if isinstance(v, (bool, np.bool_)):
            return bool(v)
        raise HsError(f"Bool expected, got {v!r}")

    # ---- pattern matching
    def match(self, p, v, env):
        k = p[0]
        if k == "pvar":
            env.d[p[1]] = v
        elif k == "pwild":
            pass
        elif k == "pbang":
            v = self.force(v)
            self.match(p[1], v, env)
        elif k == "pas":
            env.d[p[1]] = v
            self.match(p[2], v, env)
        elif k == "plit":
            v = self.force(v)
            a, b, _ = unify(v, Lit(p[1]))
            if isinstance(a, Lit):
                ok = a.q == b.q
            else:
                ok = bool(a == b)
            if not ok:
                raise MatchFail()
        elif k == "ptuple":
            v = self.force(v)
            if not isinstance(v, tuple) or len(v) != len(p[1]):
                raise HsError(f"tuple pattern against {v!r}")
            for q, x in zip(p[1], v):
                self.match(q, x, env)
        elif k == "pcon":
            name, ps = p[1], p[2]
            v = self.force(v)
            if name in ("True", "False"):
                if self.truth(v) != (name == "True"):
                    raise MatchFail()
                return
            if name in ("F#", "D#"):
                ty = F32 if name == "F#" else F64
                v = to_type(v, ty)
                if not isinstance(v, ty):
                    raise HsError(f"{name} pattern against {v!r}")
                self.match(ps[0], v, env)
                return
            if not isinstance(v, Con):
                raise HsError(f"constructor pattern {name} against {v!r}")
            if v.name != name:
                raise MatchFail()
            if len(v.args) != len(ps):
                raise HsError(f"constructor {name}: {len(v.args)} fields, pattern has {len(ps)}")
            for q, x in zip(ps, v.args):
                self.match(q, x, env)
        elif k == "pview":
            r = self.apply(self.eval(p[1], env), v)
            self.match(p[2], r, env)
        else:
            raise HsError(f"pattern {k}")

    # ---- expressions
    def eval(self, e, env):
        k = e[0]
        if k == "var":
            return env.lookup(e[1])
        if k == "lit":
            return Lit(e[1], e[2])
        if k == "app":
            return self.apply(self.eval(e[1], env), self.delay(e[2], env))
        if k == "binop":
            op = e[1]
            if op == "$":
                return self.apply(self.eval(e[2], env), self.delay(e[3], env))
            if op == "&&":
                return self.truth(self.force(self.eval(e[2], env))) and self.truth(self.force(self.eval(e[3], env)))
            if op == "||":
                return self.truth(self.force(self.eval(e[2], env))) or self.truth(self.force(self.eval(e[3], env)))
            f = env.lookup("(" + op + ")" if op in FIXITY else op)
            return self.apply_n(f, self.delay(e[2], env), self.delay(e[3], env))
        if k == "neg":
            return self.apply(env.lookup("negate"), self.delay(e[1], env))
        if k == "con":
            name = e[1]
            if name == "True":
                return True
            if name == "False":
                return False
            if name in ("F#", "D#"):
                ty = F32 if name == "F#" else F64
                return Builtin(name, 1, lambda x, ty=ty: self._typed_float(x, ty))
            return Con(name)
        if k == "tuple":
            return tuple(self.delay(x, env) for x in e[1])
        if k == "lam":
            return Closure([(e[1], [(None, e[2])], [])], len(e[1]), env)
        if k == "let":
            env2 = Env(env)
            self.bind_decls(e[1], env2)
            return self.eval(e[2], env2)
        if k == "if":
            return self.eval(e[2] if self.truth(self.force(self.eval(e[1], env))) else e[3], env)
        if k == "case":
            scrut = self.delay(e[1], env)
            for p, rhs, wh in e[2]:
                env2 = Env(env)
                try:
                    self.match(p, scrut, env2)
                except MatchFail:
                    continue
                r = self.eval_rhs(rhs, wh, env2, fall=True)
                if r is not MatchFail:
                    return r
            raise HsError("non-exhaustive case")
        if k == "rsec":      # (op e) = \x -> x op e
            f = env.lookup("(" + e[1] + ")") if e[1] in FIXITY else env.lookup(e[1])
            r = self.delay(e[2], env)
            return Builtin("rsec", 1, lambda x: self.apply_n(f, x, r))
        if k == "lsec":      # (e op) = \y -> e op y
            f = env.lookup("(" + e[1] + ")") if e[1] in FIXITY else env.lookup(e[1])
            return self.apply(f, self.delay(e[2], env))
        if k == "typed":
            v = self.force(self.eval(e[1], env))
            ty = {"Int": int, "Float": F32, "Double": F64}.get(e[2])
            if ty is None:
                raise HsError(f"type annotation {e[2]!r}")
            return to_type(v, ty)
        if k == "reccon":
            info = self.datas.get(e[1])
            if info is None or info["fields"] is None:
                raise HsError(f"{e[1]} is not a record constructor")
            if e[2] is None:            # RecordWildCards: every field from the variable of the same name
                vals = [self.force(env.lookup(fn)) for fn in info["fields"]]
            else:
                given = dict(e[2])
                vals = [self.force(self.eval(given[fn], env)) for fn in info["fields"]]
            return Con(e[1], vals)
        if k == "rhs":
            return self.eval_rhs(e[1], e[2], env)
        if k == "py":
            return e[1]()
        if k == "patsel":
            _, pat, th, nm = e
            env2 = Env()
            try:
                self.match(pat, th, env2)
            except MatchFail:
                raise HsError("irrefutable pattern failed")
            return env2.d[nm]
        raise HsError(f"expression {k}")

    def _typed_float(self, x, ty):
        x = to_type(self.force(x), ty)
        if not isinstance(x, ty):
            raise HsError(f"{ty} expected, got {x!r}")
        return x

    # ---- prelude: base primitives and the vector combinators (see the module docstring)
    def _install_prelude(self):
        g = self.glob.d
        F = self.force

        def b(name, arity):
            def deco(fn):
                g[name] = Builtin(name, arity, fn)
                return fn
            return deco

        def arith(name, fq, fnp):
            @b("(" + name + ")", 2)
            def _(x, y, fq=fq, fnp=fnp):
                x, y, ty = unify(F(x), F(y))
                if ty is None:
                    return Lit(fq(x.q, y.q), x.frac or y.frac)
                if ty is int:
                    return int(fnp(x, y))
                return ty(fnp(x, y))

        arith("+", lambda p, q: p + q, lambda p, q: p + q)
        arith("-", lambda p, q: p - q, lambda p, q: p - q)
        arith("*", lambda p, q: p * q, lambda p, q: p * q)

        @b("(/)", 2)
        def _(x, y):
            x, y, ty = unify(F(x), F(y))
            if ty is None:
                return Lit(x.q / y.q, True)
            if ty is int:
                raise HsError("(/) at Int")
            return ty(x / y)

        def cmp_op(name, fq):
            @b("(" + name + ")", 2)
            def _(x, y, fq=fq):
                x, y, ty = unify(F(x), F(y))
                if ty is None:
                    return bool(fq(x.q, y.q))
                return bool(fq(x, y))

        cmp_op("==", lambda p, q: p == q)
        cmp_op("/=", lambda p, q: p != q)
        cmp_op("<", lambda p, q: p < q)
        cmp_op("<=", lambda p, q: p <= q)
        cmp_op(">", lambda p, q: p > q)
        cmp_op(">=", lambda p, q: p >= q)

        @b("($)", 2)
        def _(f, x):
            return self.apply(f, x)

        @b("(.)", 3)
        def _(f, h, x):
            return self.apply(f, Thunk(("py", lambda: self.apply(h, x)), None))

        @b("(&&)", 2)
        def _(x, y):
            return self.truth(F(x)) and self.truth(F(y))

        @b("(||)", 2)
        def _(x, y):
            return self.truth(F(x)) or self.truth(F(y))

        @b("(^)", 2)
        def _(x0, y0):
            # GHC.Real (^) of base-4.9: repeated squaring
            x0, y0 = F(x0), to_type(F(y0), int)
            if not isinstance(y0, int):
                raise HsError("(^): Int exponent expected")
            if y0 < 0:
                raise HsError("Negative exponent")
            mul = lambda p, q: F(self.apply_n(g["(*)"], p, q))  # noqa: E731
            if y0 == 0:
                return Lit(1)

            def f(x, y):
                while y % 2 == 0:
                    x, y = mul(x, x), y // 2
                if y == 1:
                    return x
                return gg(mul(x, x), (y - 1) // 2, x)

            def gg(x, y, z):
                while True:
                    if y % 2 == 0:
                        x, y = mul(x, x), y // 2
                    elif y == 1:
                        return mul(x, z)
                    else:
                        x, y, z = mul(x, x), (y - 1) // 2, mul(x, z)
            return f(x0, y0)

        @b("negate", 1)
        def _(x):
            x = F(x)
            if isinstance(x, Lit):
                return Lit(-x.q, x.frac)
            return -x if isinstance(x, int) else type(x)(-x)

        @b("abs", 1)
        def _(x):
            x = F(x)
            if isinstance(x, Lit):
                return Lit(abs(x.q), x.frac)
            if isinstance(x, int):
                return abs(x)
            # GHC.Float (base-4.9): abs x | x == 0 = 0 | x > 0 = x | otherwise = negateFloat x
            if x == 0:
                return type(x)(0)
            return x if x > 0 else type(x)(-x)

        @b("signum", 1)
        def _(x):
            x = F(x)
            if isinstance(x, Lit):
                return Lit((x.q > 0) - (x.q < 0))
            if isinstance(x, int):
                return (x > 0) - (x < 0)
            # GHC.Float (base-4.9): signum x | x > 0 = 1 | x < 0 = negateFloat 1 | otherwise = x
            if x > 0:
                return type(x)(1)
            if x < 0:
                return type(x)(-1)
            return x

        @b("sqrt", 1)
        def _(x):
            x = F(x)
            if isinstance(x, Lit):
                # the type is not fixed yet (nrm2 over zeros only: 0.0 * sqrt 1.0): exact when the root is rational
                from math import isqrt
                p_, q_ = x.q.numerator, x.q.denominator
                if p_ >= 0 and isqrt(p_) ** 2 == p_ and isqrt(q_) ** 2 == q_:
                    return Lit(Fraction(isqrt(p_), isqrt(q_)), True)
                raise HsError("sqrt of an untyped literal")
            return type(x)(np.sqrt(x))

        def compare(x, y):
            x, y, ty = unify(F(x), F(y))
            if ty is None:
                x, y = x.q, y.q
            return Con("LT") if x < y else Con("EQ") if x == y else Con("GT")
        g["compare"] = Builtin("compare", 2, compare)

        @b("comparing", 3)
        def _(f, x, y):
            return compare(self.apply(f, x), self.apply(f, y))

        g["otherwise"] = True

        @b("return", 1)      # the Identity monad of the index streams (Util.hs:39-60)
        def _(x):
            return x

        @b("fst", 1)
        def _(p):
            return F(p)[0]

        @b("snd", 1)
        def _(p):
            return F(p)[1]

        @b("succ", 1)
        def _(x):
            return to_type(F(x), int) + 1

        # GHC.Exts
        @b("float2Double#", 1)
        def _(x):
            return F64(self._typed_float(x, F32))

        @b("double2Float#", 1)
        def _(x):
            return F32(self._typed_float(x, F64))

        # ---- Data.Vector.Storable (vector-0.11), fused
        def vec(v):
            v = F(v)
            if not isinstance(v, Vec):
                raise HsError(f"Vector expected, got {v!r}")
            return v

        def intarg(v):
            v = to_type(F(v), int)
            if not isinstance(v, int) or isinstance(v, bool):
                raise HsError(f"Int expected, got {v!r}")
            return v

        @b("V.foldl'", 3)
        def _(f, z, v):
            acc = F(z)
            for x in vec(v).iter():
                acc = F(self.apply_n(f, acc, x))
            return acc

        @b("V.map", 2)
        def _(f, v):
            v = vec(v)
            return Vec(gen=lambda: (F(self.apply(f, x)) for x in v.iter()))

        @b("V.zipWith", 3)
        def _(f, a, c):
            a, c = vec(a), vec(c)
            return Vec(gen=lambda: (F(self.apply_n(f, x, y)) for x, y in zip(a.iter(), c.iter())))

        @b("V.unsafeTake", 2)
        def _(n, v):
            n, items = intarg(n), vec(v).force()
            if n < 0 or n > len(items):
                raise HsError(f"unsafeTake {n} of a vector of {len(items)}: undefined in the reference")
            return Vec(items=items[:n])

        @b("V.unsafeIndex", 2)
        def _(v, i):
            i, items = intarg(i), vec(v).force()
            if i < 0 or i >= len(items):
                raise HsError(f"unsafeIndex {i} of a vector of {len(items)}: undefined in the reference")
            return items[i]

        @b("V.replicate", 2)
        def _(n, x):
            n = intarg(n)

            def gen():
                if n > 0:
                    xv = F(x)
                    for _ in range(n):
                        yield xv
            return Vec(gen=gen)

        @b("V.reverse", 1)
        def _(v):
            return Vec(items=list(reversed(vec(v).force())))

        @b("V.length", 1)
        def _(v):
            return len(vec(v).force())

        @b("V.unsafeUpdate_", 3)
        def _(v, ixs, ws):
            # clone of the whole destination, then the writes in stream order; the zip of the index stream
            # and the value stream ends with the shorter one
            out = list(vec(v).force())
            for i, w in zip(vec(ixs).iter(), vec(ws).iter()):
                i = intarg(i)
                if i < 0 or i >= len(out):
                    raise HsError(f"unsafeUpdate_ index {i} outside a vector of {len(out)}: undefined in the reference")
                out[i] = F(w)
            return Vec(items=out)

        @b("V.maxIndexBy", 2)
        def _(cmp, v):
            # vector-0.11 Data.Vector.Generic.maxIndexBy: foldl1' over (index, element) pairs with
            #   imax (i,x) (j,y) = case cmp x y of LT -> (j,y) ; _ -> (i,x)
            it = enumerate(vec(v).iter())
            try:
                bi, bx = next(it)
            except StopIteration:
                raise HsError("maxIndexBy: empty vector")
            for j, y in it:
                if F(self.apply_n(cmp, bx, y)).name == "LT":
                    bi, bx = j, y
            return bi

        # ---- Data.Vector.Fusion: the index streams of Util.hs
        @b("fromStream", 2)
        def _(s, size):
            return Con("Bundle", [s, size])

        @b("unstream", 1)
        def _(bundle):
            bundle = F(bundle)
            stream = F(bundle.args[0])
            if not isinstance(stream, Con) or stream.name != "Stream" or len(stream.args) != 2:
                raise HsError("unstream: Stream step state expected")
            step, s0 = stream.args

            def gen():
                s = s0
                while True:
                    r = F(self.apply(step, s))
                    if r.name == "Done":
                        return
                    if r.name == "Skip":
                        s = r.args[0]
                        continue
                    if r.name != "Yield":
                        raise HsError(f"stream step returned {r!r}")
                    yield F(r.args[0])
                    s = r.args[1]
            return Vec(gen=gen)

    # ---- the host side: marshalling numpy <-> values
    def vector(self, a):
        a = np.asarray(a)
        ty = {np.dtype(np.float32): F32, np.dtype(np.float64): F64}[a.dtype]
        return Vec(items=[ty(x) for x in a.reshape(-1)])

    def deep(self, v, ty):
        """force a result completely; literals take the type `ty`; vectors become numpy arrays"""
        v = self.force(v)
        if isinstance(v, Lit):
            return to_type(v, ty)
        if isinstance(v, tuple):
            return tuple(self.deep(x, ty) for x in v)
        if isinstance(v, Vec):
            return np.array([to_type(self.force(x), ty) for x in v.force()], dtype=ty)
        if isinstance(v, Con):
            return Con(v.name, [self.deep(x, ty) for x in v.args])
        return v

    def call(self, name, *args, ty=F32):
        def run():
            with np.errstate(all="ignore"):
                f = self.glob.lookup(name)
                return self.deep(self.apply_n(f, *args) if args else f, ty)
        return _with_big_stack(run)


def _with_big_stack(fn):
    """the reference's loops are tail calls; here they recurse -- run them on a thread with a deep stack"""
    if threading.current_thread().name == "hs_eval":
        return fn()
    box = {}

    def target():
        try:
            box["v"] = fn()
        except BaseException as ex:      # noqa: BLE001
            box["e"] = ex
    old_limit = sys.getrecursionlimit()
    sys.setrecursionlimit(1_000_000)
    old = threading.stack_size(1024 * 1024 * 1024)
    try:
        t = threading.Thread(target=target, name="hs_eval")
        t.start()
        t.join()
    finally:
        threading.stack_size(old)
        sys.setrecursionlimit(old_limit)
    if "e" in box:
        raise box["e"]
    return box["v"]


REFERENCE_SRC = "/root/reference/src/Numerical/BLAS"
REFERENCE_FILES = ("Types.hs", "Util.hs", "Single.hs")


def load_reference(src_dir: str = REFERENCE_SRC) -> Interp:
    """An interpreter with the reference's three hot-path modules loaded from where they lie."""
    it = Interp()
    import os
    for f in REFERENCE_FILES:
        it.load(os.path.join(src_dir, f))
    return it


# --------------------------------------------------------------------------------------------------
# the reference's pure call surface over numpy (same shapes as oracle/oracle.py, so tests can swap them)
# --------------------------------------------------------------------------------------------------
MGR_FIELDS = {"FLAGNEG2": (), "FLAGNEG1": ("d1", "d2", "x1", "h11", "h12", "h21", "h22"),
              "FLAG0": ("d1", "d2", "x1", "h12", "h21"), "FLAG1": ("d1", "d2", "x1", "h11", "h22")}
MGR_FLAG = {"FLAGNEG2": -2, "FLAGNEG1": -1, "FLAG0": 0, "FLAG1": 1}


class Reference:
    """`Numerical.BLAS.Single` evaluated from its source.  Vectors in and out are numpy arrays (float32 selects
    the Float instance, float64 the Double instance); a ModGivensRot is (constructor name, {field: value})."""

    def __init__(self, src_dir: str = REFERENCE_SRC):
        self.it = load_reference(src_dir)
        # the record declaration itself fixes the positional order of the fields (Types.hs:15-20)
        for con, fields in MGR_FIELDS.items():
            got = self.it.datas[con]["fields"] or ()
            if tuple(got) != fields:
                raise HsError(f"{con}: fields {got} in the reference, {fields} expected")

    def _v(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype not in (np.float32, np.float64):
            raise TypeError("float32 or float64 vectors")
        return self.it.vector(a), (F64 if a.dtype == np.float64 else F32)

    def _call(self, name, ty, *args):
        return self.it.call(name, *args, ty=ty)

    def dot(self, n, x, incx, y, incy):
        (vx, ty), (vy, _) = self._v(x), self._v(y)
        return self._call("dot", ty, int(n), vx, int(incx), vy, int(incy))

    def sdsdot(self, n, sb, x, incx, y, incy):
        (vx, _), (vy, _) = self._v(np.asarray(x, np.float32)), self._v(np.asarray(y, np.float32))
        return self._call("sdsdot", F32, int(n), F32(sb), vx, int(incx), vy, int(incy))

    def asum(self, n, x, incx):
        vx, ty = self._v(x)
        return self._call("asum", ty, int(n), vx, int(incx))

    def nrm2(self, n, x, incx):
        vx, ty = self._v(x)
        return self._call("nrm2", ty, int(n), vx, int(incx))

    def iamax(self, n, x, incx):
        vx, ty = self._v(x)
        return int(self._call("iamax", ty, int(n), vx, int(incx)))

    def scal(self, n, a, x, incx):
        vx, ty = self._v(x)
        return self._call("scal", ty, int(n), ty(a), vx, int(incx))

    def copy(self, n, x, incx, y, incy):
        (vx, ty), (vy, _) = self._v(x), self._v(y)
        return self._call("copy", ty, int(n), vx, int(incx), vy, int(incy))

    def swap(self, n, x, incx, y, incy):
        (vx, ty), (vy, _) = self._v(x), self._v(y)
        return self._call("swap", ty, int(n), vx, int(incx), vy, int(incy))

    def axpy(self, n, a, x, incx, y, incy):
        (vx, ty), (vy, _) = self._v(x), self._v(y)
        return self._call("axpy", ty, int(n), ty(a), vx, int(incx), vy, int(incy))

    def rot(self, n, x, incx, y, incy, c, s):
        (vx, ty), (vy, _) = self._v(x), self._v(y)
        return self._call("rot", ty, int(n), vx, int(incx), vy, int(incy), ty(c), ty(s))

    def rotm(self, flags, n, x, incx, y, incy):
        (vx, ty), (vy, _) = self._v(x), self._v(y)
        con, fields = flags
        arg = Con(con, [ty(fields[f]) for f in MGR_FIELDS[con]])
        return self._call("rotm", ty, arg, int(n), vx, int(incx), vy, int(incy))

    def rotg(self, sa, sb, ty=F32):
        r = self._call("rotg", ty, ty(sa), ty(sb))
        return tuple(r.args[0])

    def rotmg(self, sd1, sd2, sx1, sy1, ty=F32):
        r = self._call("rotmg", ty, ty(sd1), ty(sd2), ty(sx1), ty(sy1))
        return r.name, dict(zip(MGR_FIELDS[r.name], r.args))
