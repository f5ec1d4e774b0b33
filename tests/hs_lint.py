"""A syntax / scope / arity lint for the Haskell sources under hs/ -- there is no GHC in this image, so the binding layer a
maintainer of the reference would compile has never met a compiler.  This is what can be checked without
one (tests/test_haskell_shim.py):

  * the files lex, lay out (Haskell 2010 layout rule) and PARSE as Haskell modules -- the parser is oracle/hs_eval.py's,
    extended here to whole modules (imports, foreign imports, classes, instances, do blocks, lists, strings, type
    annotations, arbitrary operators); it is first run over the reference's own modules, which GHC accepts;
  * every identifier a module uses is defined in it, imported by name, a member of a module it imports whole (member
    lists below, from the packages' documentation), or in the Prelude;
  * every use of a foreign import or class method applies it to as many arguments as its signature has (or to none: the
    function is passed on), and every instance defines every method of its class.

It is NOT a type checker."""
from oracle import hs_eval as H


OPS = dict(H.FIXITY)
OPS.update({">>=": (1, "l"), ">>": (1, "l"), "=<<": (1, "r"), "<$>": (4, "l"), "<$": (4, "l"), "$>": (4, "l"), "<*>": (4, "l"),
            "*>": (4, "l"), "<*": (4, "l"), "++": (5, "r"), ":": (5, "r"), "!!": (9, "l"), "$!": (0, "r"), "<>": (6, "r"),
            ".&.": (7, "l"), ".|.": (5, "l"), "!": (9, "l"), "<|>": (3, "l"), "&": (1, "l")})
BACKTICK = {"div": (7, "l"), "mod": (7, "l"), "quot": (7, "l"), "rem": (7, "l"), "seq": (0, "r"), "elem": (4, "n"), "notElem": (4, "n")}
TERMINATORS_KW = {"then", "else", "of", "in", "where", "deriving"}


class FullParser(H.Parser):
    def __init__(self, toks, fname="?"):
        super().__init__(toks, fname)
        self.imports, self.foreign, self.classes, self.instances, self.types = [], {}, {}, [], set()
        self.sigs = {}

    # ---- module level -------------------------------------------------------------------------
    def parse_module(self):
        decls, datas = [], {}
        self.exports = None
        if self.is_("kw", "module"):
            self.next()
            self.next()                                    # module name
            if self.is_("special", "("):
                self.exports = self.paren_names()
            self.expect("kw", "where")
        else:
            raise H.HsError(f"{self.fname}: module header expected")
        self.expect("special", "{")
        while not self.is_("special", "}") and not self.is_("eof"):
            if self.accept("special", ";"):
                continue
            t = self.peek()
            if t.kind == "kw" and t.val == "import":
                self.parse_import()
            elif t.kind == "varid" and t.val == "foreign" and self.is_("kw", "import", 1):
                self.parse_foreign()
            elif t.kind == "kw" and t.val in ("type", "newtype"):
                self.next()
                self.types.add(self.peek().val)
                if t.val == "newtype":
                    self.parse_newtype(datas)
                else:
                    self.skip_decl()
            elif t.kind == "kw" and t.val == "data":
                self.types.add(self.peek(1).val)
                k = 0
                has_eq = False
                while not (self.is_("special", ";", k) or self.is_("special", "}", k) or self.is_("eof", None, k)):
                    if self.is_("op", "=", k):
                        has_eq = True
                        break
                    k += 1
                if has_eq:
                    self.parse_data(datas)
                else:
                    self.skip_decl()                     # `data LbkCtx`: no constructors
            elif t.kind == "kw" and t.val == "class":
                self.parse_class(decls)
            elif t.kind == "kw" and t.val == "instance":
                self.parse_instance()
            else:
                d = self.parse_decl()
                if d is not None:
                    decls.append(d)
        return decls, datas

    def paren_names(self):
        """names inside ( ... ): [(name, sub-list or None)]"""
        self.expect("special", "(")
        names = []
        while not self.is_("special", ")"):
            t = self.next()
            if t.kind in ("varid", "conid"):
                sub = None
                if self.is_("special", "("):
                    sub = [n for n, _ in self.paren_names()] or [".."]
                names.append((t.val, sub))
            elif t.kind == "special" and t.val == "(":           # an operator: (<+>)
                names.append(("(" + self.next().val + ")", None))
                self.expect("special", ")")
            elif t.kind == "op" and t.val == "..":
                names.append(("..", None))
            elif t.kind == "kw" and t.val in ("module", "type"):
                pass
        self.next()
        return names

    def parse_import(self):
        self.expect("kw", "import")
        qualified = self.accept("kw", "qualified") is not None
        mod = self.expect("conid").val
        alias = mod
        if self.accept("kw", "as"):
            alias = self.expect("conid").val
        hiding = self.accept("kw", "hiding") is not None
        names = None
        if self.is_("special", "("):
            names = self.paren_names()
        self.imports.append({"module": mod, "qualified": qualified, "alias": alias, "hiding": hiding, "names": names})
        self.skip_decl()

    def type_arity(self):
        """skips a type up to the end of the declaration; returns the number of top-level arrows"""
        d, arrows = 0, 0
        while True:
            t = self.peek()
            if t.kind == "eof":
                break
            if t.kind == "special":
                if t.val in "([{":
                    d += 1
                elif t.val in ")]}":
                    if d == 0:
                        break
                    d -= 1
                elif t.val == ";" and d == 0:
                    break
            if d == 0 and t.kind == "op" and t.val == "->":
                arrows += 1
            if d == 0 and t.kind == "op" and t.val == "=>":
                arrows = 0                                 # what came before was the context
            self.next()
        return arrows

    def parse_foreign(self):
        self.next(); self.next()                           # foreign import
        conv = self.expect("varid").val
        if conv not in ("ccall", "capi", "stdcall"):
            raise H.HsError(f"{self.fname}: calling convention {conv}")
        if self.is_("varid", "unsafe") or self.is_("varid", "safe") or self.is_("varid", "interruptible"):
            self.next()
        ent = self.expect("str").val
        name = self.expect("varid").val
        self.expect("op", "::")
        self.foreign[name] = {"entity": ent, "arity": self.type_arity()}

    def parse_newtype(self, datas):
        while not self.is_("op", "="):
            self.next()
        self.next()
        con = self.expect("conid").val
        fields = None
        if self.is_("special", "{"):
            self.next()
            fields = [self.expect("varid").val]
            self.expect("op", "::")
            while not self.is_("special", "}"):
                self.next()
            self.next()
        datas[con] = {"fields": fields, "arity": 1}
        self.skip_decl()

    def parse_class(self, decls):
        self.expect("kw", "class")
        head = []
        while not (self.is_("kw", "where") or self.is_("special", ";") or self.is_("special", "}")):
            head.append(self.next())
        names = [t.val for t in head if t.kind == "conid"]
        cname = names[-1] if not any(t.kind == "op" and t.val == "=>" for t in head) else \
            [t.val for t in head[[i for i, t in enumerate(head) if t.kind == "op" and t.val == "=>"][-1]:] if t.kind == "conid"][0]
        self.types.add(cname)
        methods, defaults = {}, set()
        if self.accept("kw", "where"):
            self.expect("special", "{")
            while not self.is_("special", "}"):
                if self.accept("special", ";"):
                    continue
                if self.is_("varid"):
                    k, ns = 1, [self.peek().val]
                    while self.is_("special", ",", k) and self.is_("varid", None, k + 1):
                        ns.append(self.peek(k + 1).val)
                        k += 2
                    if self.is_("op", "::", k):
                        self.i += k + 1
                        ar = self.type_arity()
                        for n in ns:
                            methods[n] = ar
                        continue
                d = self.parse_decl()
                if d is not None:
                    decls.append(d)
                    if d[0] == "fun":
                        defaults.add(d[1])
            self.next()
        self.classes[cname] = {"methods": methods, "defaults": defaults}

    def parse_instance(self):
        self.expect("kw", "instance")
        head = []
        while not (self.is_("kw", "where") or self.is_("special", ";") or self.is_("special", "}") or self.is_("eof")):
            head.append(self.next())
        arrows = [i for i, t in enumerate(head) if t.kind == "op" and t.val == "=>"]
        body = head[arrows[-1] + 1:] if arrows else head
        cname = body[0].val
        defs = []
        if self.accept("kw", "where"):
            defs = self.parse_block()
        self.instances.append({"class": cname, "head": " ".join(str(t.val) for t in body[1:]), "decls": defs})

    def parse_decl(self):
        # signatures are remembered (arity) before the base class skips them
        if self.is_("varid"):
            k, ns = 1, [self.peek().val]
            while self.is_("special", ",", k) and self.is_("varid", None, k + 1):
                ns.append(self.peek(k + 1).val)
                k += 2
            if self.is_("op", "::", k):
                self.i += k + 1
                ar = self.type_arity()
                for n in ns:
                    self.sigs[n] = ar
                return None
        return super().parse_decl()

    # ---- expressions ---------------------------------------------------------------------------
    def cur_op(self):
        t = self.peek()
        if t.kind == "op" and t.val not in H.RESERVED_OPS | {"!"} or (t.kind == "op" and t.val in (":", "!") and False):
            p, a = OPS.get(t.val, (9, "l"))
            return t.val, p, a, 1
        if t.kind == "op" and t.val == ":":
            return ":", 5, "r", 1
        if t.kind == "special" and t.val == "`" and self.is_("special", "`", 2):
            p, a = BACKTICK.get(self.peek(1).val.split(".")[-1], (9, "l"))
            return self.peek(1).val, p, a, 3
        return None

    def parse_exp(self):
        e = self.parse_opexp(0)
        if self.is_("op", "::"):                          # a type annotation: skip the type
            self.next()
            d = 0
            while True:
                t = self.peek()
                if t.kind == "eof" or (d == 0 and ((t.kind == "special" and t.val in ";,)]}") or (t.kind == "kw" and t.val in TERMINATORS_KW)
                                                  or (t.kind == "op" and t.val in ("=", "|", "<-")))):
                    break
                if t.kind == "special" and t.val in "([":
                    d += 1
                elif t.kind == "special" and t.val in ")]":
                    d -= 1
                self.next()
            e = ("typed", e, "?")
        return e

    def parse_exp10(self):
        if self.is_("kw", "do"):
            self.next()
            self.expect("special", "{")
            stmts = []
            while not self.is_("special", "}"):
                if self.accept("special", ";"):
                    continue
                stmts.append(self.parse_stmt())
            self.next()
            if not stmts or stmts[-1][0] != "exp":
                raise H.HsError(f"{self.fname}: the last statement of a do block must be an expression (line {self.peek().line})")
            return ("do", stmts)
        return super().parse_exp10()

    def parse_stmt(self):
        if self.is_("kw", "let"):
            self.next()
            decls = self.parse_block()
            if self.accept("kw", "in"):                   # it was an expression after all
                return ("exp", ("let", decls, self.parse_exp()))
            return ("let", decls)
        d, k, bind = 0, 0, False
        while True:
            t = self.peek(k)
            if t.kind == "eof":
                break
            if t.kind == "special" and t.val in "([{":
                d += 1
            elif t.kind == "special" and t.val in ")]}":
                if d == 0:
                    break
                d -= 1
            elif d == 0 and t.kind == "special" and t.val == ";":
                break
            elif d == 0 and t.kind == "op" and t.val == "<-":
                bind = True
                break
            elif d == 0 and t.kind == "op" and t.val in ("->", "="):
                break
            k += 1
        if bind:
            p = self.parse_pat()
            self.expect("op", "<-")
            return ("bind", p, self.parse_exp())
        return ("exp", self.parse_exp())

    def aexp_start(self):
        t = self.peek()
        return t.kind in ("varid", "conid", "num", "str", "char") or (t.kind == "special" and t.val in "([")

    def parse_aexp(self):
        t = self.peek()
        if t.kind in ("str", "char"):
            self.next()
            return ("lit", 0, False)
        if t.kind == "special" and t.val == "[":
            self.next()
            items = []
            while not self.is_("special", "]"):
                items.append(self.parse_exp())
                if not (self.accept("special", ",") or self.accept("op", "..") or self.accept("op", "|") or self.accept("op", "<-")):
                    break
            self.expect("special", "]")
            return ("list", items)
        if t.kind == "special" and t.val == "(" and self.is_("special", ",", 1):       # (,) (,,)
            self.next()
            while self.accept("special", ","):
                pass
            self.expect("special", ")")
            return ("var", "(,)")
        e = super().parse_aexp()
        while self.is_("special", "{") and self.peek().extra != "virtual" and e[0] in ("var", "app"):      # record update
            self.next()
            fs = []
            while not self.is_("special", "}"):
                fn = self.expect("varid").val
                self.expect("op", "=")
                fs.append((fn, self.parse_exp()))
                self.accept("special", ",")
            self.next()
            e = ("recupd", e, fs)
        return e

    # ---- patterns ------------------------------------------------------------------------------
    def apat_start(self):
        t = self.peek()
        return t.kind in ("varid", "conid", "num", "str", "char") or (t.kind == "special" and t.val in "([") or (t.kind == "op" and t.val in ("!", "~"))

    def record_pat(self, con):
        """Con{..} / Con{} / Con{f = p, ...} after the constructor name"""
        self.expect("special", "{")
        fs = []
        while not self.is_("special", "}"):
            if self.accept("op", ".."):
                fs.append(("..", None))
            else:
                fn = self.expect("varid").val
                fs.append((fn, self.parse_pat() if self.accept("op", "=") else None))
            self.accept("special", ",")
        self.next()
        return ("precord", con, fs)

    def parse_pat(self):
        if self.is_("conid") and self.is_("special", "{", 1) and self.peek(1).extra != "virtual":
            p = self.record_pat(self.next().val)
        else:
            p = super().parse_pat()
        if self.is_("op", ":"):                            # x : xs
            self.next()
            return ("pcon", ":", [p, self.parse_pat()])
        return p

    def parse_apat(self):
        t = self.peek()
        if t.kind == "conid" and self.is_("special", "{", 1) and self.peek(1).extra != "virtual":
            return self.record_pat(self.next().val)
        if t.kind == "varid" and self.is_("op", "@", 1) and self.is_("conid", None, 2) and self.is_("special", "{", 3):
            self.next(); self.next()
            return ("pas", t.val, self.record_pat(self.next().val))
        if t.kind in ("str", "char"):
            self.next()
            return ("pwild",)
        if t.kind == "op" and t.val == "~":
            self.next()
            return self.parse_apat()
        if t.kind == "special" and t.val == "[":
            self.next()
            ps = []
            while not self.is_("special", "]"):
                ps.append(self.parse_pat())
                self.accept("special", ",")
            self.next()
            return ("pcon", "[]", ps)
        return super().parse_apat()


def parse_file(path):
    with open(path) as f:
        src = f.read()
    p = FullParser(H.layout(H.lex(src)), path)
    decls, datas = p.parse_module()
    if not p.is_("special", "}") and not p.is_("eof"):
        raise H.HsError(f"{path}: trailing tokens at {p.peek()!r}")
    return p, decls, datas


# ---- scope ------------------------------------------------------------------------------------------
PRELUDE = set("""
abs all and any asTypeOf atan2 break ceiling compare concat concatMap const cos curry cycle div divMod drop dropWhile either elem
error even exp fail filter flip floor fmap foldl foldr foldr1 foldl1 fromEnum fromInteger fromIntegral fromRational fst gcd head id init
isNaN isInfinite iterate last lcm length lines log lookup map mapM mapM_ max maxBound maximum maybe min minBound minimum mod negate not notElem
null odd or otherwise pi pred print product properFraction pure putStr putStrLn quot quotRem read realToFrac recip rem repeat replicate
return reverse round scanl seq sequence sequence_ show signum sin snd span splitAt sqrt subtract succ sum tail take takeWhile tan toEnum
toInteger toRational truncate uncurry undefined unlines until unwords unzip unzip3 userError words zip zip3 zipWith zipWith3 mempty mappend
mconcat traverse foldMap sequenceA floatDigits
True False Nothing Just Left Right LT EQ GT IO Int Integer Float Double Bool Char String Maybe Either Ordering Word Show Eq Ord Num Fractional
Floating Real RealFrac RealFloat Integral Enum Bounded Functor Applicative Monad Monoid Read
""".split())
# members of modules imported WITHOUT an import list (or qualified), as documented for base / vector
MEMBERS = {
    "Foreign.C.Types": "CInt CUInt CLong CULong CLLong CSize CChar CFloat CDouble CUChar CShort CUShort".split(),
    "Foreign.ForeignPtr": "ForeignPtr FinalizerPtr newForeignPtr newForeignPtr_ withForeignPtr finalizeForeignPtr touchForeignPtr castForeignPtr mallocForeignPtr mallocForeignPtrArray mallocForeignPtrBytes addForeignPtrFinalizer plusForeignPtr".split(),
    "Foreign.Ptr": "Ptr FunPtr nullPtr castPtr plusPtr minusPtr alignPtr nullFunPtr castFunPtr castFunPtrToPtr castPtrToFunPtr freeHaskellFunPtr IntPtr WordPtr ptrToIntPtr intPtrToPtr".split(),
    "Data.Vector.Storable": "Vector Storable MVector length null head last fromList toList unsafeWith unsafeFromForeignPtr0 unsafeToForeignPtr0 unsafeFreeze unsafeThaw freeze thaw map zipWith foldl' replicate generate empty singleton unsafeIndex unsafeTake reverse sum maximum minimum concat take drop slice (!) unsafeUpdate_ maxIndexBy".split(),
    "Data.Vector.Storable.Mutable": "IOVector MVector STVector new unsafeNew replicate length unsafeWith read write unsafeRead unsafeWrite slice".split(),
    "Control.DeepSeq": "NFData rnf deepseq force ($!!)".split(),
    "GHC.Generics": ["Generic"],
    # what hs/test/DeviceMain.hs imports whole, as the reference's test/Main.hs does
    "Test.Tasty": "TestTree TestName defaultMain testGroup localOption".split(),
    "Test.Tasty.QuickCheck": "testProperty Gen Property forAll choose elements suchThat ioProperty frequency vectorOf oneof listOf property (==>) (===) QuickCheckTests".split(),
    # the reference's own modules: src/Numerical/BLAS/Single.hs:41-61 (export list), test/Gen.hs
    "Numerical.BLAS.Single": ("asum sasum dasum nrm2 snrm2 dnrm2 dot sdot ddot sdsdot scal sscal dscal rot srot drot rotg srotg drotg rotm srotm drotm "
                              "rotmg srotmg drotmg axpy saxpy daxpy copy scopy dcopy swap sswap dswap iamax isamax idamax").split(),
    "Gen": "genNiceFloat genNiceDouble genNVector genFloat genDouble genEveryday genNegative".split(),
}


def names_used(e, bound, out):
    """free variables / constructors of an expression, declaration or pattern tree (the parser's tuples)"""
    if isinstance(e, list):
        for x in e:
            names_used(x, bound, out)
        return
    if not isinstance(e, tuple) or not e:
        return
    k = e[0]
    if k in ("var", "con"):
        if e[1] not in bound:
            out.add(e[1])
    elif k == "binop":
        if e[1] not in bound:
            out.add(e[1] if e[1] not in OPS and not e[1].startswith("(") else "(" + e[1] + ")")
        names_used(e[2], bound, out); names_used(e[3], bound, out)
    elif k in ("lsec", "rsec"):
        out.add(e[1] if e[1] not in OPS else "(" + e[1] + ")")
        names_used(e[2], bound, out)
    elif k == "lam":
        names_used(e[2], bound | pat_vars(e[1]), out)
        pat_cons(e[1], out, bound)
    elif k == "let":
        b = bound | decl_names(e[1])
        decls_used(e[1], b, out)
        names_used(e[2], b, out)
    elif k == "case":
        names_used(e[1], bound, out)
        for p, rhs, wh in e[2]:
            b = bound | pat_vars([p]) | decl_names(wh)
            pat_cons([p], out, bound)
            decls_used(wh, b, out)
            for g, x in rhs:
                names_used(g, b, out); names_used(x, b, out)
    elif k == "do":
        b = set(bound)
        for st in e[1]:
            if st[0] == "bind":
                names_used(st[2], b, out)
                pat_cons([st[1]], out, b)
                b |= pat_vars([st[1]])
            elif st[0] == "let":
                b |= decl_names(st[1])
                decls_used(st[1], b, out)
            else:
                names_used(st[1], b, out)
    elif k == "reccon":
        out.add(e[1])
        for _, x in (e[2] or []):
            names_used(x, bound, out)
    elif k == "recupd":
        names_used(e[1], bound, out)
        for f, x in e[2]:
            if f not in bound:
                out.add(f)
            names_used(x, bound, out)
    elif k in ("lit",):
        pass
    else:
        for x in e[1:]:
            if isinstance(x, (tuple, list)):
                names_used(x, bound, out)


def pat_vars(ps):
    out = set()
    for p in ps:
        k = p[0]
        if k == "pvar":
            out.add(p[1])
        elif k == "pas":
            out.add(p[1]); out |= pat_vars([p[2]])
        elif k == "pbang":
            out |= pat_vars([p[1]])
        elif k == "ptuple":
            out |= pat_vars(p[1])
        elif k == "pcon":
            out |= pat_vars(p[2])
        elif k == "pview":
            out |= pat_vars([p[2]])
        elif k == "precord":
            for f, q in p[2]:
                if q is not None:
                    out |= pat_vars([q])
                elif f != "..":
                    out.add(f)                      # NamedFieldPuns
    return out


def pat_cons(ps, out, bound):
    for p in ps:
        k = p[0]
        if k == "pcon":
            if p[1] not in ("[]", ":"):
                out.add(p[1])
            pat_cons(p[2], out, bound)
        elif k in ("pas",):
            pat_cons([p[2]], out, bound)
        elif k == "pbang":
            pat_cons([p[1]], out, bound)
        elif k == "ptuple":
            pat_cons(p[1], out, bound)
        elif k == "pview":
            names_used(p[1], bound, out); pat_cons([p[2]], out, bound)
        elif k == "precord":
            out.add(p[1])
            pat_cons([q for _, q in p[2] if q is not None], out, bound)


def decl_names(decls):
    out = set()
    for d in decls:
        if d[0] == "fun":
            out.add(d[1])
        else:
            out |= pat_vars([d[1]])
    return out


def decls_used(decls, bound, out):
    for d in decls:
        if d[0] == "fun":
            _, name, pats, rhs, wh = d
            b = bound | pat_vars(pats) | decl_names(wh)
            pat_cons(pats, out, bound)
        else:
            _, pat, rhs, wh = d
            b = bound | decl_names(wh)
            pat_cons([pat], out, bound)
        decls_used(wh, b, out)
        for g, x in rhs:
            names_used(g, b, out); names_used(x, b, out)


def applications(e, visit):
    """calls visit(head_name, nargs) for every application spine whose head is a variable"""
    if isinstance(e, list):
        for x in e:
            applications(x, visit)
        return
    if not isinstance(e, tuple) or not e:
        return
    if e[0] == "app" or (e[0] == "binop" and e[1] == "$"):           # f a b  and  f a $ b  are both applications of f
        n, f = 0, e
        while f[0] == "app" or (f[0] == "binop" and f[1] == "$"):
            applications(f[2] if f[0] == "app" else f[3], visit)
            n += 1
            f = f[1] if f[0] == "app" else f[2]
        if f[0] == "var":
            visit(f[1], n)
        else:
            applications(f, visit)
        return
    if e[0] == "var":
        visit(e[1], 0)
        return
    for x in e[1:]:
        if isinstance(x, (tuple, list)):
            applications(x, visit)
