"""The reference's Python surface, kept name for name, on top of the device engine.

    ExactMatch  <- reference SMEM/ExactMatch.py:7-199
    LUT         <- reference SMEM/LUT.py:6-63
    RMI         <- reference SMEM/RMI.py:4-86
    RMI_LUT     <- reference SMEM/RMI_LUT.py:10-198
    SMEM        <- reference SMEM/SMEM.py:8-484

Same constructor arguments, attributes, return conventions (tuple (lo, hi) of 0-based inclusive
rows, int -1 on a miss, dict[str -> (lo, hi)] for SMEM sets, 1-based positions) and the same
cwd-relative "data/" layout.  Every search runs on the GPU through libgenie_smem; each class also
has a *_batch method, which is what a pipeline should call (one launch for many reads).
Differences, all deliberate:
  * create_fm_index uses an O(n) suffix sorter and writes data/<name>-FM.npz (binary) instead of a
    pretty-printed JSON; load_fm_index reads either that or the reference's <name>-FM.json;
  * non-ACGT characters raise BaseError (a KeyError and a ValueError) before any launch;
  * get_smems_rmi loads rmi_file.pkl once per SMEM object, not once per call (SMEM.py:207).
"""
import json
import os
import pickle
import random
from os import path

import numpy as np

from . import engine as eng
from . import _capi as capi

_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def _check_bases(seq):
    for ch in seq:
        if ch not in _CODE:
            raise eng.BaseError(ch)


class _FmIndexView(dict):
    """fm_index of the reference (ExactMatch.py:29-30), materialised lazily from the host index."""

    def __init__(self, host):
        super().__init__()
        self._host = host
        dict.__setitem__(self, "ref_size", host.n_rows)
        dict.__setitem__(self, "count_dic", host.count_dic())

    def __missing__(self, key):
        if key in ("suffix_array", "bwt_array"):
            sa, bwt = self._host.export()
            dict.__setitem__(self, "suffix_array", sa)
            dict.__setitem__(self, "bwt_array", list(bwt.decode()))
            return dict.__getitem__(self, key)
        if key == "occurance_matrix":
            bwt = np.frombuffer("".join(self["bwt_array"]).encode(), dtype=np.uint8)
            om = {c: np.cumsum(bwt == ord(c)).astype(np.int64) for c in "$ACGT"}
            dict.__setitem__(self, key, om)
            return om
        raise KeyError(key)

    def __bool__(self):
        return True


class _DeviceBuiltIndex:
    """Adapter with HostIndex's read-only surface (info, n_rows, export, count_dic) over an index that was built ON THE
    GPU (engine.DeviceIndex.build_on_device): what create_fm_index uses for large references."""

    def __init__(self, dev_index, text):
        self.dev, self._text = dev_index, text
        self.info = dev_index.info
        self.n_rows, self.n_bases = dev_index.n_rows, dev_index.n_bases

    def export(self):
        sa = self.dev.suffix_array_host().copy()
        t = np.frombuffer((self._text + "$").encode(), np.uint8)
        bwt = t[(sa.astype(np.int64) - 2) % len(t)]          # the character before each suffix ('$' before the whole text)
        return sa, bwt.tobytes()

    def count_dic(self):
        return self.dev.count_dic()


# references at least this long are indexed on the GPU by create_fm_index / from_text (host SA-IS below)
DEVICE_BUILD_MIN_BASES = 1 << 22


class ExactMatch:
    def __init__(self, reference_sequence_file: str, query_sequence_file: str = None, device="cuda"):
        self.ref_seq_file = reference_sequence_file
        self.ref_sequence = None
        self.ref_size = None
        self.query_sequence = None
        self.fm_file = self.ref_seq_file.split(".")[0] + "-FM.json"
        self.fm_index = {}
        self._npz = self.ref_seq_file.split(".")[0] + "-FM.npz"
        self._device = device
        self._host = None
        self._dev = None
        if query_sequence_file is not None:
            self.load_query(query_sequence_file)

    # ------------------------------------------------------------------ index
    def load_ref_sequence(self):
        with open(path.join("data", self.ref_seq_file), "r") as f:
            f.readline()
            self.ref_sequence = "".join(line.strip() for line in f)
        self.ref_sequence += "$"
        self.ref_size = len(self.ref_sequence)

    def create_fm_index(self, reference_json=False):
        """Writes data/<ref>-FM.npz (binary suffix array).  reference_json=True also writes data/<ref>-FM.json in the
        reference's own schema (ExactMatch.py:29-33), loadable by the reference's load_fm_index / RMI_LUT / train.py."""
        self.load_ref_sequence()
        self._build(self.ref_sequence[:-1])
        sa, _ = self._host.export()
        np.savez(path.join("data", self._npz), suffix_array=sa, ref_size=np.int64(self.ref_size))
        if reference_json:
            self.export_reference_json()

    def export_reference_json(self, file=None):
        """The index in the reference's on-disk schema: {bwt_array, suffix_array, occurance_matrix, count_dic, ref_size}
        exactly as ExactMatch.create_fm_index dumps it (ExactMatch.py:29-33; inclusive occurrence counts per
        row, 1-based suffix array).  O(5n) Python ints: for references of the reference's own scale."""
        fm = self.fm_index
        obj = {"bwt_array": list(fm["bwt_array"]), "suffix_array": [int(x) for x in fm["suffix_array"]],
               "occurance_matrix": {c: [int(x) for x in v] for c, v in fm["occurance_matrix"].items()},
               "count_dic": dict(fm["count_dic"]), "ref_size": int(fm["ref_size"])}
        file = file or path.join("data", self.fm_file)
        with open(file, "w+") as f:
            f.write(json.dumps(obj, indent=4, sort_keys=True))
        return file

    def load_fm_index(self):
        """data/<ref>-FM.npz or the reference's data/<ref>-FM.json (ExactMatch.py:35-41); when both exist the newer file
        wins.  The loaded suffix array must fit the text (ref_size rows), else ValueError."""
        if self.ref_sequence is None:
            self.load_ref_sequence()
        npz, js = path.join("data", self._npz), path.join("data", self.fm_file)
        have = [p for p in (npz, js) if path.exists(p)]
        if not have:
            raise FileNotFoundError("No FM index file found. Run ExactMatch.createFMIndex to create an FM index.")
        src = max(have, key=path.getmtime)
        if src == npz:
            sa = np.load(npz)["suffix_array"]
        else:
            with open(js, "r") as f:
                fm = json.load(f)
            fm = fm.get("fm_index", fm)
            if int(fm.get("ref_size", len(fm["suffix_array"]))) != len(self.ref_sequence):
                raise ValueError(f"{js}: ref_size does not match {self.ref_seq_file}")
            sa = np.asarray(fm["suffix_array"], dtype=np.uint32)
        if len(sa) != len(self.ref_sequence):
            raise ValueError(f"{src}: suffix array length {len(sa)} does not match the reference ({len(self.ref_sequence)} rows)")
        self._set_host(eng.HostIndex.from_arrays(self.ref_sequence[:-1], sa))

    @classmethod
    def from_text(cls, text, suffix_array=None, device="cuda", name="memory.fa", builder="auto"):
        """Build directly from a string (no files): used by benchmarks and tests.  builder: "host" (SA-IS), "device"
        (gsm_index_build_device) or "auto" (device for references of DEVICE_BUILD_MIN_BASES bases and more)."""
        m = cls(name, device=device)
        m.ref_sequence = text + "$"
        m.ref_size = len(text) + 1
        if suffix_array is not None:
            m._set_host(eng.HostIndex.from_arrays(text, suffix_array))
        else:
            m._build(text, builder)
        return m

    def _build(self, text, builder="auto"):
        """The index of `text`: on the GPU for large references (seconds at 10^9 bases), host SA-IS otherwise.
        Replaces the n^2 rotation sort of ExactMatch.create_fm_index (reference ExactMatch.py:52-58)."""
        import torch
        on_device = builder == "device" or (builder == "auto" and len(text) >= DEVICE_BUILD_MIN_BASES and torch.cuda.is_available())
        if not on_device:
            self._set_host(eng.HostIndex.build(text))
            return
        _check_bases(text) if len(text) < 1 << 16 else None       # large texts: the device packer validates
        dev = eng.DeviceIndex.build_on_device(text, self._device)
        self._set_host(_DeviceBuiltIndex(dev, text))
        self._dev = dev

    def _set_host(self, host):
        self._host = host
        self._dev = None
        self.ref_size = host.n_rows
        self.fm_index = _FmIndexView(host)

    @property
    def device_index(self):
        if self._host is None:
            self.load_fm_index()
        if self._dev is None:
            self._dev = eng.DeviceIndex(self._host, self._device)
        return self._dev

    # ------------------------------------------------------------------ queries
    def load_query(self, query_seq_file):
        with open(path.join("data", query_seq_file), "r") as f:
            self.query_sequence = "".join(line.strip() for line in f)

    def create_query(self, query_size, query_output_file=None):
        if self.ref_sequence is None:
            self.load_ref_sequence()
        start = random.randint(0, self.ref_size - query_size - 1)
        query = self.ref_sequence[start:start + query_size]
        if query_output_file is not None:
            with open(path.join("data", query_output_file), "w+") as f:
                f.write(">query from '" + self.ref_seq_file + "' zero-index location: " + str(start) + "\n")
                f.write("\n".join(query[i:i + 50] for i in range(0, len(query), 50)))
        self.query_sequence = query
        return query

    # ------------------------------------------------------------------ search (GPU)
    def exact_match_back_prop_batch(self, queries):
        """-> (lo, cnt) uint32 arrays; cnt == 0 means the reference returns -1."""
        for q in queries:
            _check_bases(q)
        return eng.backsearch_batch(self.device_index, eng.ReadBatch.from_strings(list(queries)))

    def exact_match_back_prop(self, query_seq: str):
        lo, cnt = self.exact_match_back_prop_batch([query_seq])
        return -1 if cnt[0] == 0 else (int(lo[0]), int(lo[0]) + int(cnt[0]) - 1)

    def exact_match_back_prop_add_one(self, char, prev_suffix_tuple):
        _check_bases(char)
        lo, cnt = eng.add_one_batch(self.device_index, [_CODE[char]], [prev_suffix_tuple[0]],
                                    [prev_suffix_tuple[1] - prev_suffix_tuple[0] + 1])
        return -1 if cnt[0] == 0 else (int(lo[0]), int(lo[0]) + int(cnt[0]) - 1)

    def exact_match(self, query_seq: str = None):
        if query_seq is None:
            if self.query_sequence is None:
                print("No query sequence. Either input sequence, load_file, or generate.")
                return
            query_seq = self.query_sequence
        start, end = self.exact_match_back_prop(query_seq)      # TypeError on a miss, as in the reference
        return sorted(self.get_positions(start, end))

    def get_position(self, suffix_array_index):
        return int(self.device_index.locate([suffix_array_index % self.ref_size])[0])

    def get_positions(self, suffix_start, suffix_end):
        rows = np.arange(suffix_start, suffix_end + 1, dtype=np.int64) % self.ref_size
        return [int(x) for x in self.device_index.locate(rows)]


class _LutView:
    """LUT.lut of the reference: str(code) -> [(lo, hi), positions] (LUT.py:33-35), read through
    to the dense device table."""

    def __init__(self, lut):
        self._l = lut

    def _entry(self, key):
        code = int(key)
        if code < 0 or code >= (1 << (2 * self._l.lut_size)):
            return None
        e = self._l.table[2 * code:2 * code + 2].cpu().numpy().view(np.uint32)
        return None if e[1] == 0 else (int(e[0]), int(e[0]) + int(e[1]) - 1)

    def __contains__(self, key):
        return self._entry(key) is not None

    def __getitem__(self, key):
        t = self._entry(key)
        if t is None:
            raise KeyError(key)
        return [t, self._l.matcher.get_positions(t[0], t[1])]

    def __len__(self):
        return int((self._l.table[1::2] != 0).sum().item())


class LUT:
    def __init__(self, matcher: ExactMatch):
        self.matcher = matcher
        if self.matcher.ref_sequence is None:
            self.matcher.load_ref_sequence()
        self.lut = None
        self.lut_size = None
        self.table = None     # device: 4^K x (lo, cnt) int32

    def generate_lut(self, size):
        self.lut_size = int(size)
        self.table = eng.lut_build(self.matcher.device_index, self.lut_size)
        self.lut = _LutView(self)

    @staticmethod
    def convert_seq_to_num(sequence):
        v = 0
        for ch in sequence:
            v = v << 2 | _CODE[ch]
        return v

    def _file(self):
        return path.join("data", self.matcher.ref_seq_file.split(".")[0] + "-LUT.json")

    def save_lut(self):
        if self.lut is None:
            raise RuntimeError("LUT has not been created yet.")
        t = self.table.cpu().numpy().view(np.uint32).reshape(-1, 2)
        out = {}
        for code in np.nonzero(t[:, 1])[0]:
            lo, cnt = int(t[code, 0]), int(t[code, 1])
            out[str(int(code))] = [[lo, lo + cnt - 1], self.matcher.get_positions(lo, lo + cnt - 1)]
        with open(self._file(), "w") as f:
            f.write(json.dumps({"lut": out, "lut_size": self.lut_size}, indent=4, sort_keys=True))

    def load_lut(self):
        """<ref>-LUT.json as LUT.save_lut of the reference writes it (LUT.py:50-63).  The dense device table is derived
        from the index for the file's lut_size; the file's own intervals are then checked against it (ValueError if the
        file belongs to another reference)."""
        with open(self._file(), "r") as f:
            obj = json.load(f)
        self.generate_lut(obj["lut_size"])
        entries = obj.get("lut", {})
        if entries:
            t = self.table.cpu().numpy().view(np.uint32).reshape(-1, 2)
            codes = np.fromiter((int(k) for k in entries), np.int64, len(entries))
            lo = np.fromiter((int(v[0][0]) for v in entries.values()), np.int64, len(entries))
            hi = np.fromiter((int(v[0][1]) for v in entries.values()), np.int64, len(entries))
            if len(entries) != int((t[:, 1] != 0).sum()) or not (np.array_equal(t[codes, 0], lo) and np.array_equal(t[codes, 0].astype(np.int64) + t[codes, 1] - 1, hi)):
                raise ValueError(f"{self._file()} does not describe this reference")


class RMI:
    """Recursive model index of linear models (reference SMEM/RMI.py).  fit() is host-side numpy
    (closed-form least squares per bucket); predict() reproduces sklearn's fl(fl(x*coef)+intercept)."""

    def __init__(self, experts=[100, 1000]):
        self.experts = list(experts)
        self.level_sizes = [1] + self.experts
        self.coef = None
        self.intercept = None

    @classmethod
    def from_params(cls, level_sizes, coef, intercept):
        r = cls(list(level_sizes[1:]))
        r.coef = np.ascontiguousarray(coef, np.float64)
        r.intercept = np.ascontiguousarray(intercept, np.float64)
        return r

    def fit(self, x, y, vectorised=None):
        x = np.asarray(x, np.float64).reshape(-1)
        y = np.asarray(y, np.float64).reshape(-1)
        if vectorised is None:
            vectorised = sum(self.level_sizes) > 8192 or x.shape[0] > 2_000_000
        if vectorised:
            return self._fit_vectorised(x, y)
        levels = self.experts + [1]
        buckets = [np.arange(x.shape[0])]
        coef, icpt = [], []
        root = None
        for scale in levels:
            nxt = [[] for _ in range(scale)]
            allocated = 0.0
            for pts in buckets:
                if len(pts) == 0:                       # RMI.py:24-26: empty buckets alias the root model
                    coef.append(root[0]); icpt.append(root[1])
                    continue
                cx, cy = x[pts], y[pts]
                if scale == 1:
                    target = cy
                else:
                    span = cy.max() - cy.min()
                    if span == 0:
                        budget = 1
                    else:
                        cy = (cy - cy.min()) / span
                        budget = len(pts) * scale / len(x)
                    target = cy * budget + allocated
                    allocated += budget
                xm, tm = cx.mean(), target.mean()
                var = ((cx - xm) ** 2).sum()
                a = 0.0 if var == 0 else float(((cx - xm) * (target - tm)).sum() / var)
                b = float(tm - a * xm)
                if root is None:
                    root = (a, b)
                coef.append(a); icpt.append(b)
                pred = cx * a + b
                route = np.clip(np.trunc(pred), 0, scale - 1).astype(np.int64)
                order = np.argsort(route, kind="stable")
                bounds = np.searchsorted(route[order], np.arange(scale + 1))
                for k in np.nonzero(np.diff(bounds))[0]:
                    nxt[k].append(pts[order[bounds[k]:bounds[k + 1]]])
            buckets = [np.concatenate(b) if b else np.empty(0, np.int64) for b in nxt]
        self.coef = np.asarray(coef, np.float64)
        self.intercept = np.asarray(icpt, np.float64)
        return self

    def _fit_vectorised(self, x, y):
        """The same training rule (reference SMEM/RMI.py:11-50: per-bucket min-max normalised targets, proportional
        budgets, closed-form line per bucket, routing by truncated prediction) with every per-bucket loop replaced by
        segment reductions, so 10^6 leaf models over 10^7..10^8 keys train in seconds.  Sums run in
        a different order than fit()'s per-bucket numpy calls, so parameters agree to rounding, not bit for bit --
        any parameter set is a valid input of the search path."""
        n = x.shape[0]
        levels = self.experts + [1]
        bid = np.zeros(n, np.int64)
        nb = 1
        coef, icpt = [], []
        root = None
        for scale in levels:
            cnt = np.bincount(bid, minlength=nb)
            ne = cnt > 0
            order = np.argsort(bid, kind="stable")
            starts = (np.cumsum(cnt) - cnt)[ne]
            ys = y[order]
            ymin = np.zeros(nb); ymax = np.zeros(nb)
            ymin[ne] = np.minimum.reduceat(ys, starts)
            ymax[ne] = np.maximum.reduceat(ys, starts)
            span = ymax - ymin
            if scale == 1:
                target = y
            else:
                budget = np.where(span == 0, 1.0, cnt * scale / n)
                budget[~ne] = 0.0
                allocated = np.cumsum(budget) - budget
                flat = span[bid] == 0
                norm = np.where(flat, y, (y - ymin[bid]) / np.where(span[bid] == 0, 1.0, span[bid]))
                target = norm * budget[bid] + allocated[bid]
            safe = np.maximum(cnt, 1)
            xm = np.bincount(bid, x, minlength=nb) / safe
            tm = np.bincount(bid, target, minlength=nb) / safe
            dx = x - xm[bid]
            var = np.bincount(bid, dx * dx, minlength=nb)
            cov = np.bincount(bid, dx * (target - tm[bid]), minlength=nb)
            a = np.where(var == 0, 0.0, cov / np.where(var == 0, 1.0, var))
            b = tm - a * xm
            if root is None:
                root = (float(a[0]), float(b[0]))
            a[~ne] = root[0]
            b[~ne] = root[1]
            coef.append(a); icpt.append(b)
            pred = x * a[bid] + b[bid]
            bid = np.clip(np.trunc(pred), 0, scale - 1).astype(np.int64)
            nb = scale
        self.coef = np.concatenate(coef).astype(np.float64)
        self.intercept = np.concatenate(icpt).astype(np.float64)
        return self

    def predict(self, x):
        x = np.asarray(x, np.float64).reshape(-1)
        off = np.concatenate([[0], np.cumsum(self.level_sizes)])[:-1]
        model = np.zeros(x.shape[0], np.int64)
        p = np.zeros(x.shape[0])
        for lvl, scale in enumerate(self.experts + [1]):
            k = off[lvl] + model
            p = x * self.coef[k] + self.intercept[k]
            model = np.clip(np.trunc(p), 0, scale - 1).astype(np.int64)
        return p

    def dump(self, filename, final_scale):
        sizes = np.asarray(self.level_sizes[1:], dtype="int32")
        w = np.stack([self.coef, self.intercept], axis=1).astype(np.float64)
        n_last = self.level_sizes[-1]
        w[-n_last:] *= final_scale
        with open(filename, "wb+") as f:
            f.write(sizes.tobytes())
            f.write(w.reshape(-1).astype("float32").tobytes())

    # reference pickles hold sklearn LinearRegression objects in .models (RMI.py:8,49)
    def __setstate__(self, state):
        self.__dict__.update(state)
        if "models" in state and state.get("coef") is None:
            self.level_sizes = [len(l) for l in state["models"]]
            self.coef = np.asarray([float(m.coef_[0]) for l in state["models"] for m in l], np.float64)
            self.intercept = np.asarray([float(m.intercept_) for l in state["models"] for m in l], np.float64)
            self.__dict__.pop("models", None)
            self.__dict__.pop("all_buckets", None)


def _reference_rmi_object(rmi):
    """An object that pickles as the reference's RMI.RMI (module "RMI"): experts, models[level][k] = LinearRegression."""
    import sys
    import types
    from sklearn.linear_model import LinearRegression
    mod = sys.modules.get("RMI")
    if mod is None or not hasattr(mod, "RMI"):
        mod = types.ModuleType("RMI")
        mod.RMI = type("RMI", (), {"__module__": "RMI"})
        sys.modules["RMI"] = mod
    obj = mod.RMI.__new__(mod.RMI)
    models, k = [], 0
    for n in rmi.level_sizes:
        level = []
        for _ in range(n):
            m = LinearRegression()
            m.coef_ = np.asarray([rmi.coef[k]], np.float64)
            m.intercept_ = np.float64(rmi.intercept[k])
            m.n_features_in_ = 1
            level.append(m)
            k += 1
        models.append(level)
    obj.__dict__.update({"experts": list(rmi.experts), "models": models})
    return obj


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "RMI" and name == "RMI":
            return RMI
        return super().find_class(module, name)


class RMI_LUT:
    def __init__(self, RMI_structure, LUT_size, data_file, matcher: ExactMatch = None):
        self.nucleo = dict(_CODE)
        self.structure = RMI_structure
        self.prediction_size = LUT_size
        self.data_file = data_file
        if matcher is None:
            matcher = ExactMatch(data_file)
            matcher.load_fm_index()
        self.matcher = matcher
        if matcher.ref_sequence is None:
            matcher.load_ref_sequence()
        self.ref_seq = matcher.ref_sequence[:-1]
        self.ref_seq_size = matcher.ref_size - 1
        self.rmi = RMI(RMI_structure)
        self._params = None

    @property
    def suffix_array(self):
        return self.matcher.fm_index["suffix_array"]

    def train_RMI(self):
        """RMI_LUT.py:36-50: keys = K-prefix codes of all suffixes long enough, targets = their rows."""
        K = self.prediction_size
        sa = np.asarray(self.suffix_array, np.int64)
        rows = np.nonzero(sa - 1 + K <= self.ref_seq_size)[0]
        codes = np.frombuffer(self.ref_seq.encode(), np.uint8)
        lut = np.zeros(256, np.int64)
        for k, v in _CODE.items():
            lut[ord(k)] = v
        c = lut[codes]
        start = sa[rows] - 1
        key = np.zeros(len(rows), np.int64)
        for j in range(K):
            key = (key << 2) | c[start + j]
        self.rmi.fit(key, rows)
        self._params = None

    @property
    def params(self):
        if self._params is None:
            self._params = eng.RmiParams(self.prediction_size, self.rmi.level_sizes, self.rmi.coef, self.rmi.intercept,
                                         self.matcher.device_index.device).build_probe_table(self.matcher.device_index)
            if self.matcher.device_index.n_rows > self.prediction_size:
                self._params.build_none_rows(self.matcher.device_index)
        return self._params

    def _encode(self, query, encoded):
        if encoded:
            return int(query)
        _check_bases(query[: self.prediction_size])
        return LUT.convert_seq_to_num(query[: self.prediction_size])

    def rmi_predict(self, query, encoded=False):
        return self.rmi.predict(np.asarray([self._encode(query, encoded)]).reshape(-1, 1))

    def get_suffix_rmi_batch(self, codes):
        return eng.rmi_lookup_batch(self.matcher.device_index, self.params, codes)

    def get_suffix_rmi(self, query, encoded=False):
        pred, lo, hi, st = self.get_suffix_rmi_batch([self._encode(query, encoded)])
        if st[0] != capi.READ_OK:
            raise RecursionError("RMI last-mile search left the table (the reference raises here too)")
        return int(lo[0]), int(hi[0])

    def save(self, file, reference_format=False):
        """Pickle [structure, prediction_size, data_file, rmi] (RMI_LUT.py:186-190).  reference_format=True writes what the
        reference's own RMI_LUT.load can unpickle: an object of class RMI.RMI whose .models are sklearn LinearRegression
        instances carrying these coefficients (needs scikit-learn; RMI.py:8,49,52-69)."""
        rmi = self.rmi
        if reference_format:
            rmi = _reference_rmi_object(self.rmi)
        with open(file, "wb") as f:
            pickle.dump([self.structure, self.prediction_size, self.data_file, rmi], f)

    @staticmethod
    def load(file, matcher: ExactMatch = None):
        with open(file, "rb") as f:
            structure, K, data_file, rmi = _RefUnpickler(f).load()
        r = RMI_LUT(structure, K, data_file, matcher)
        r.rmi = rmi
        return r


class SMEM:
    def __init__(self, matcher: ExactMatch):
        self.matcher = matcher
        self.lut = LUT(self.matcher)
        try:
            self.lut.load_lut()          # SMEM.py:11-12 loads <ref>-LUT.json eagerly
        except FileNotFoundError:
            pass                          # generate with smem.lut.generate_lut(K)
        self.rmi_lut = None
        self._engine = None

    # ------------------------------------------------------------------ batched entry points
    def _run(self, method, queries, **kw):
        queries = list(queries)
        for q in queries:
            _check_bases(q)
        if not queries:
            return eng.SmemResult(np.zeros(0, eng.RECORD_DTYPE), np.zeros(1, np.int64), np.zeros(0, np.uint8), 0)
        idx = self.matcher.device_index
        if not idx.all_bases_present:
            raise ValueError("the reference text must contain all of A, C, G, T (SURVEY 8c parity domain)")
        batch = eng.ReadBatch.from_strings(queries)
        if self._engine is None or self._engine.max_reads < batch.n or self._engine.max_len < batch.max_len:
            self._engine = eng.Engine(idx, max(batch.n, 1024), max(batch.max_len, 160), mems_per_read=64, recs_per_read=48)
        return self._engine.run(method, batch, **kw)

    def get_SMEMS_batch(self, queries, minimum_length):
        return self._run(capi.METHOD_BWA, queries, min_len=minimum_length)

    def get_smems_lut_batch(self, queries):
        if self.lut.table is None:
            raise RuntimeError("LUT has not been created yet.")
        return self._run(capi.METHOD_LUT, queries, K=self.lut.lut_size, lut=self.lut.table)

    def get_smems_rmi_batch(self, queries):
        if self.rmi_lut is None:
            self.rmi_lut = RMI_LUT.load("rmi_file.pkl", self.matcher)
        return self._run(capi.METHOD_RMI, queries, rmi=self.rmi_lut.params)

    @staticmethod
    def _dict(query, recs):
        out = {}
        for r in recs:
            out[query[int(r["qstart"]):int(r["qend"])]] = (int(r["sa_lo"]), int(r["sa_hi"]))
        return out

    # ------------------------------------------------------------------ reference-shaped calls
    def get_SMEMS(self, query, minimum_length):
        return self._dict(query, self.get_SMEMS_batch([query], minimum_length).for_read(0))

    def get_smems_lut(self, query):
        return self._dict(query, self.get_smems_lut_batch([query]).for_read(0))

    def get_smems_rmi(self, query):
        res = self.get_smems_rmi_batch([query])
        if res.status[0] == capi.READ_REF_RAISES:
            raise RecursionError("RMI last-mile search left the table (the reference raises here too)")
        return self._dict(query, res.for_read(0))

    def get_suffix_index(self, query):
        return self.matcher.exact_match_back_prop(query)

    def forward_extension(self, query, start_index, largest="", suffix_tuple=None):
        """SMEM.py:425-443: all prefixes are searched in ONE batched launch."""
        found = {}
        if suffix_tuple is not None:
            found[largest] = suffix_tuple
        cands = [largest + query[start_index:i] for i in range(start_index + 1, len(query) + 1)]
        if not cands:
            return found, largest
        lo, cnt = self.matcher.exact_match_back_prop_batch(cands)
        cur = largest
        for k, s in enumerate(cands):
            cur = s
            if cnt[k] == 0:
                return found, s[:-1]
            found[s] = (int(lo[k]), int(lo[k]) + int(cnt[k]) - 1)
        return found, cur

    def backward_extension(self, query, start_index, forward_matches):
        """SMEM.py:389-423 with the left extensions of all keys searched in one batched launch."""
        best, best_t, best_end = "", None, -1
        longest_fwd = ""
        keys = list(forward_matches)
        cands = [query[i:start_index] + key for key in keys for i in range(start_index - 1, -1, -1)]
        lo, cnt = self.matcher.exact_match_back_prop_batch(cands) if cands else ([], [])
        pos = 0
        for key in keys:
            if len(key) > len(longest_fwd):
                longest_fwd = key
            alive = True
            for i in range(start_index - 1, -1, -1):
                if alive:
                    if cnt[pos] == 0:
                        alive = False
                    elif len(cands[pos]) > len(best):
                        best, best_t, best_end = cands[pos], (int(lo[pos]), int(lo[pos]) + int(cnt[pos]) - 1), start_index + len(key)
                pos += 1
        if len(longest_fwd) > len(best):
            best, best_t, best_end = longest_fwd, forward_matches[longest_fwd], start_index + len(longest_fwd)
        return best, best_t, best_end

    def get_SMEM_at_index(self, query, start_index):
        fwd = self.forward_extension(query, start_index)
        back = self.backward_extension(query, start_index, fwd[0])
        if len(fwd[1]) > len(back[0]):
            return [fwd[1], fwd[0][fwd[1]], len(fwd[1]) + start_index]
        return [back[0], back[1], back[2]]

    @staticmethod
    def check_sequential(list1, list2):
        s2 = set(list2)
        return any(x + 1 in s2 for x in list1)


def create_random_query(query_size):
    return "".join(random.choice(["A", "G", "C", "T"]) for _ in range(query_size))


def create_query_from_ref(ref_seq, query_size):
    ref_size = len(ref_seq)
    query = ""
    while len(query) < query_size:
        position = random.randint(0, ref_size)
        size = random.randint(1, 30)
        if size + position > ref_size:
            continue
        query += ref_seq[position:position + size]
    return query[:query_size]
