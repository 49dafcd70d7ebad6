"""`:destination => :iostream`: the CSV output format of Klara's BasicContParamIOStream
(src/iostreams/ParameterIOStreams/BasicContParamIOStream.jl:47-82 file naming, :152-159 one comma-joined line
per saved state; wired by initialize_output, src/jobs/jobs.jl:193-201; README.md:117-146).

Files: `<filepath>/value.<suffix>`, `logtarget.<suffix>`, `gradlogtarget.<suffix>`, `diagnosticvalues.<suffix>`
for the monitored fields.  A batched job (one BasicMCJob per chain in the reference, each with its own
outopts[:filepath]) writes chain c under `<filepath>/chain<c>/` (c = global 1-based chain index); a job built from
a single vector writes directly under `<filepath>` exactly like the reference.

Numbers are printed the way Julia's `string(::Float64)` prints them (shortest round-trip digits, fixed notation for
-4 < decimal point position <= 6, otherwise d.ddde<exp>), Bool as true/false, so the files are byte-compatible
with what the reference writes for the same values.
"""
import os

import numpy as np

FIELDS = ("value", "logtarget", "gradlogtarget")


def julia_float(x):
    """string(x::Float64) of Julia 0.6 (base/grisu: shortest digits; exponential form iff pt <= -4 or pt > 6)"""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if np.signbit(x) else "0.0"
    sign = "-" if x < 0 else ""
    mant, exp = ("%r" % abs(x)).lower(), 0
    if "e" in mant:
        mant, e = mant.split("e")
        exp = int(e)
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    pt = len(ip.lstrip("0")) + exp if ip.strip("0") else exp - (len(fp) - len(fp.lstrip("0")))
    digits = digits.rstrip("0") or "0"
    n = len(digits)
    if pt <= -4 or pt > 6:                                   # d.ddde<exp>
        return "%s%s.%se%d" % (sign, digits[0], digits[1:] or "0", pt - 1)
    if pt <= 0:
        return "%s0.%s%s" % (sign, "0" * (-pt), digits)
    if pt >= n:
        return "%s%s%s.0" % (sign, digits, "0" * (pt - n))
    return "%s%s.%s" % (sign, digits[:pt], digits[pt:])


def _line(values):
    return ",".join(julia_float(v) for v in np.atleast_1d(values))


# how string() prints the entries of state.diagnosticvalues (an Array{Any} in the reference): :accept is a Bool, :ndoublings and
# :na are Ints, :a is a Float64 (src/samplers/iterate/NUTS.jl:384-399)
_DIAG_FORMAT = {"accept": lambda v: "true" if v else "false", "ndoublings": lambda v: "%d" % int(v), "na": lambda v: "%d" % int(v),
                "a": julia_float}


def _diag_token(tok):
    return 1.0 if tok == "true" else 0.0 if tok == "false" else float(tok.replace("Inf", "inf").replace("NaN", "nan"))


class BasicContParamIOStream:
    """Writer / reader of one chain's CSV files."""

    def __init__(self, size, n, monitor=("value",), filepath="", filesuffix="csv", diagnostickeys=(), mode="w"):
        self.size, self.n = size, n
        self.diagnostickeys = list(diagnostickeys)
        self.filepath, self.filesuffix = filepath, filesuffix
        self.names = {f: os.path.join(filepath, "%s.%s" % (f, filesuffix)) for f in FIELDS if f in monitor}
        if self.diagnostickeys:
            self.names["diagnosticvalues"] = os.path.join(filepath, "diagnosticvalues.%s" % filesuffix)
        if filepath and mode == "w":
            os.makedirs(filepath, exist_ok=True)
        self.mode = mode

    def write_nstate(self, value=None, logtarget=None, gradlogtarget=None, diagnosticvalues=None):
        """one line per saved state in every monitored file (write(iostream, state) called npost times)"""
        data = {"value": value, "logtarget": logtarget, "gradlogtarget": gradlogtarget}
        for f, path in self.names.items():
            with open(path, self.mode) as fh:
                if f == "diagnosticvalues":
                    keys = self.diagnostickeys
                    dv = np.asarray(diagnosticvalues)
                    if len(keys) > 1:                    # (nkeys, n) as output(job) holds it -> one line per saved state
                        for row in dv.T:
                            fh.write(",".join(_DIAG_FORMAT.get(k, julia_float)(v) for k, v in zip(keys, row)) + "\n")
                    else:
                        fmt = _DIAG_FORMAT.get(keys[0], julia_float) if keys else _DIAG_FORMAT["accept"]
                        for row in np.atleast_1d(dv):
                            fh.write(",".join(fmt(b) for b in np.atleast_1d(row)) + "\n")
                else:
                    for row in data[f]:
                        fh.write(_line(row) + "\n")

    def read(self, dtype=np.float64):
        """read(iostream, T): back into arrays (value (n, size), logtarget (n,), diagnosticvalues (n,) bool)"""
        out = {}
        for f, path in self.names.items():
            if f == "diagnosticvalues":
                with open(path) as fh:
                    rows = [ln.strip().split(",") for ln in fh if ln.strip()]
                if self.diagnostickeys in ([], ["accept"]):
                    out[f] = np.array([[tok == "true" for tok in r] for r in rows])
                else:                                    # (n, nkeys) float64: true / false read back as 1 / 0
                    out[f] = np.array([[_diag_token(tok) for tok in r] for r in rows])
            else:
                a = np.loadtxt(path, delimiter=",", dtype=dtype, ndmin=2)
                out[f] = a[:, 0] if f == "logtarget" else a
        return out


def write_job_output(job, out):
    """save the fetched NState of a BasicMCJob as CSV files following outopts"""
    oo = job.outopts
    fp, sfx = oo.get("filepath", ""), oo.get("filesuffix", "csv")
    diag = list(oo["diagnostics"])
    streams = []
    nchains = 1 if job.single else job.nchains
    for c in range(nchains):
        path = fp if job.single else os.path.join(fp, "chain%d" % (job.cfg.chain_offset + c + 1))
        st = BasicContParamIOStream(job.dim, job.range.npoststeps, oo["monitor"], path, sfx, diag)
        pick = (lambda a: a) if job.single else (lambda a: None if a is None else a[c])
        st.write_nstate(pick(out.value), pick(out.logtarget), pick(out.gradlogtarget), pick(out.diagnosticvalues))
        streams.append(st)
    return streams[0] if job.single else streams
