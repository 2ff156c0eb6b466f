"""Host-side mirror of the reference's compile-time type rules (Python side of the boundary).

The C++23 header `include/QuBLAS.h` of this repo does the same lowering in `consteval` code; this
module is its Python twin, used by the ctypes binding, the tests and bench.py.

Reference (file:line in /root/reference/include/QuBLAS.h):
  Qu defaults 8/8/signed/TRN::TCPL/SAT::TCPL              :2355-2359
  MulMerger / AddMerger / FullPrec                         :3079-3140
  TFComplexMul tag routing (incl. its quirks)              :3510-3535  
  Reducer level selection and result type                  :4903-5001
"""
from dataclasses import dataclass, replace
from typing import Dict, Optional, Sequence

# QuMode ids (QuBLAS.h:1988-1998) and OfMode ids (:2212-2223)
RND_POS_INF, RND_NEG_INF, RND_ZERO, RND_INF, RND_CONV, TRN_TCPL, TRN_SMGN = range(7)
SAT_TCPL, SAT_ZERO, SAT_SMGN, WRP_TCPL = range(4)
QU_NAMES = ["RND::POS_INF", "RND::NEG_INF", "RND::ZERO", "RND::INF", "RND::CONV", "TRN::TCPL", "TRN::SMGN"]
OF_NAMES = ["SAT::TCPL", "SAT::ZERO", "SAT::SMGN", "WRP::TCPL"]

TREE_VECTOR, TREE_VARIADIC = 0, 1


@dataclass(frozen=True)
class QType:
    """Qu<intBits<i>, fracBits<f>, isSigned<s>, QuMode<q>, OfMode<o>> (QuBLAS.h:2368-2385)."""
    i: int = 8
    f: int = 8
    s: bool = True
    q: int = TRN_TCPL
    o: int = SAT_TCPL

    def __post_init__(self):
        if self.i + self.f < 0:                      # static_assert at QuBLAS.h:2372
            raise ValueError("intBits + fracBits must be non-negative")

    @property
    def width(self) -> int:
        """Stored width: a sign bit is always present (QuBLAS.h:2384-2385)."""
        return 1 + self.i + self.f

    @property
    def word_bytes(self) -> int:
        """The reference's storage word: int32 for width <= 32, else int64 (QuBLAS.h:353)."""
        return 4 if self.width <= 32 else 8

    @property
    def raw_min(self) -> int:
        return -(1 << (self.width - 1))

    @property
    def raw_max(self) -> int:
        return (1 << (self.width - 1)) - 1

    @staticmethod
    def from_dict(d) -> "QType":
        return QType(int(d["i"]), int(d["f"]), bool(d["s"]), int(d["q"]), int(d["o"]))

    def to_dict(self):
        return {"i": self.i, "f": self.f, "s": self.s, "q": self.q, "o": self.o}


def _tags_of(t) -> Optional[dict]:
    """Merger arguments: None, a dict of partial tags (i/f/s/q/o/full), or a full QType
    (MergerArgsWrapper unpacks a Qu type into its five tags, QuBLAS.h:3097-3099)."""
    if t is None:
        return {}
    if isinstance(t, QType):
        return t.to_dict()
    return dict(t)


def _merge(a: QType, b: QType, tags, is_mul: bool) -> QType:
    tags = _tags_of(tags)
    full = bool(tags.get("full"))
    if is_mul:      # MulMerger, QuBLAS.h:3114-3115
        di = a.i + b.i if full else max(a.i, b.i)
        df = a.f + b.f if full else max(a.f, b.f)
    else:           # AddMerger, QuBLAS.h:3133-3134
        di = max(a.i, b.i) + 1 if full else max(a.i, b.i)
        df = max(a.f, b.f)
    dq = a.q if a.q == b.q else TRN_TCPL            # :3111 / :3130
    do = a.o if a.o == b.o else SAT_TCPL            # :3112 / :3131
    return QType(tags.get("i", di), tags.get("f", df), bool(tags.get("s", a.s or b.s)),
                 tags.get("q", dq), tags.get("o", do))


def mul_merge(a: QType, b: QType, tags=None) -> QType:
    return _merge(a, b, tags, True)


def add_merge(a: QType, b: QType, tags=None) -> QType:
    return _merge(a, b, tags, False)


TF_TAGS = ("abT", "cdT", "baT", "abcT", "cdbT", "badT", "ABT", "BCT")


def resolve_tf(x_re: QType, x_im: QType, y_re: QType, y_im: QType, tags: Optional[Dict] = None) -> Dict[str, QType]:
    """Types of the eight TFComplexMul nodes for x = a+bi, y = c+di (QuBLAS.h:3510-3535).

    Reproduces the reference's routing: product B=(c+d)*b is governed by badT and C=(b-a)*d by
    cdbT (:3525-3526); baT is ignored because `subbaType` is extracted as ::type, not ::type::list
    (:3515), so b-a gets the default AddMerger.  A top-level FullPrec reaches every node whose own
    tag is missing (their default list is the whole TF argument pack) except b-a.
    """
    tags = dict(tags or {})
    top_full = {"full": True} if tags.get("FullPrec") or tags.get("full") else None
    g = lambda k: tags[k] if k in tags else top_full
    a, b, c, d = x_re, x_im, y_re, y_im
    r = {}
    r["ab"] = add_merge(a, b, g("abT"))
    r["cd"] = add_merge(c, d, g("cdT"))
    r["ba"] = add_merge(b, a, None)
    r["abc"] = mul_merge(r["ab"], c, g("abcT"))
    r["bad"] = mul_merge(r["cd"], b, g("badT"))
    r["cdb"] = mul_merge(r["ba"], d, g("cdbT"))
    r["AB"] = add_merge(r["abc"], r["bad"], g("ABT"))
    r["BC"] = add_merge(r["bad"], r["cdb"], g("BCT"))
    return r


BCM_TAGS = ("acT", "bdT", "adT", "bcT", "acbdT", "adbcT")


def resolve_bcm(x_re: QType, x_im: QType, y_re: QType, y_im: QType, tags: Optional[Dict] = None) -> Dict[str, QType]:
    """Types of the six BasicComplexMul nodes (QuBLAS.h:3426-3446): re = ac - bd, im = ad + bc.  This is the
    product `Qmul(c1, c2)` uses when no method is named (:3422-3424)."""
    tags = dict(tags or {})
    top_full = {"full": True} if tags.get("FullPrec") or tags.get("full") else None
    g = lambda k: tags[k] if k in tags else top_full
    a, b, c, d = x_re, x_im, y_re, y_im
    r = {}
    r["ac"] = mul_merge(a, c, g("acT"))
    r["bd"] = mul_merge(b, d, g("bdT"))
    r["ad"] = mul_merge(a, d, g("adT"))
    r["bc"] = mul_merge(b, c, g("bcT"))
    r["acbd"] = add_merge(r["ac"], r["bd"], g("acbdT"))
    r["adbc"] = add_merge(r["ad"], r["bc"], g("adbcT"))
    return r


def resolve_rc(real_first: bool, r: QType, c_re: QType, c_im: QType, tags=None) -> Dict[str, QType]:
    """Product types of real (x) complex Qmul (QuBLAS.h:3600-3637).  `tags` is the merger argument pack:
    {"realT": t, "imagT": t} per component, a (t_real, t_imag) pair (two positional types, :3618-3619), or one
    QType / partial-tag dict / None that serves both components.  Keys follow qb_bcm_types: real operand first
    (x = a, y = c+di): ac, ad; complex operand first (x = a+bi, y = c): ac, bc."""
    if isinstance(tags, dict) and ("realT" in tags or "imagT" in tags):
        rest = {k: v for k, v in tags.items() if k not in ("realT", "imagT")} or None
        t_re, t_im = tags.get("realT", rest), tags.get("imagT", rest)
    elif isinstance(tags, (tuple, list)):
        t_re, t_im = tags
    else:
        t_re = t_im = tags
    if real_first:
        return {"ac": mul_merge(r, c_re, t_re), "ad": mul_merge(r, c_im, t_im)}
    return {"ac": mul_merge(c_re, r, t_re), "bc": mul_merge(c_im, r, t_im)}


def level_type(levels: Sequence[QType], layer: int, a: QType, b: QType) -> QType:
    """TypeAt<min(layer, n-1)> (QuBLAS.h:4913); no levels -> default AddMerger (:4917-4921)."""
    if not levels:
        return add_merge(a, b)
    return levels[min(layer, len(levels) - 1)]


def tree_out_type(n: int, elem: QType, levels: Sequence[QType], form: int = TREE_VECTOR, layer: int = 0) -> QType:
    """Result type of Qreduce over n elements (vector form :4960-4990, variadic form :4924-4957)."""
    if n < 1:
        raise ValueError("empty reduction")
    if form == TREE_VARIADIC:
        if n == 1:
            return elem
        lt = level_type(levels, layer, elem, elem)
        r = tree_out_type(n // 2, lt, levels, form, layer + 1)
        if n & 1:
            r = level_type(levels, layer, r, elem)
        return r
    t = elem
    while n > 1:
        t = level_type(levels, layer, t, t)
        n = (n + 1) // 2
        layer += 1
    return t
