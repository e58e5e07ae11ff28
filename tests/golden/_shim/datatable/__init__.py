"""Minimal pandas-backed stand-in for the third-party `datatable` package.

TEST INFRASTRUCTURE ONLY.  `datatable==1.0.0` (pinned by the reference's
requirements.txt:1) has no wheel in this image and there is no network, so the
reference (recommender.py:13, main.py:3, generator.py:3) cannot be imported
as-is.  `tests/golden/make_golden.py` puts this directory first on `sys.path`
so the UNMODIFIED reference modules run and produce the golden vectors.

Only the surface the reference touches is provided:
  dt.fread(path) / dt.fread(path, header=False) / dt.fread(path, columns=[...])
  dt.Frame(dict), dt.str64, Frame.shape/.names/.to_numpy()/.to_pandas()
  frame[:] = dt.str64 ; del frame[:, [names]] ; frame[:, update(**exprs)]
  f[name] == None, ifelse(cond, a, b), update(**kwargs)
"""
import numpy as np
import pandas as pd


class _StrType:
    pass


class _Col:
    def __init__(self, name):
        self.name = name

    def __eq__(self, other):
        assert other is None
        return _IsNA(self.name)

    __hash__ = None


class _IsNA:
    def __init__(self, name):
        self.name = name


class _F:
    def __getitem__(self, name):
        return _Col(name)


class _IfElse:
    def __init__(self, cond, a, b):
        self.cond, self.a, self.b = cond, a, b


class _Update:
    def __init__(self, exprs):
        self.exprs = exprs


def ifelse(cond, a, b):
    return _IfElse(cond, a, b)


def update(**kwargs):
    return _Update(kwargs)


f = _F()


class Frame:
    def __init__(self, data=None):
        if isinstance(data, pd.DataFrame):
            self._df = data
        elif data is None:
            self._df = pd.DataFrame()
        else:
            self._df = pd.DataFrame({k: np.asarray(v) for k, v in dict(data).items()})

    @property
    def shape(self):
        return self._df.shape

    @property
    def names(self):
        return tuple(self._df.columns)

    def to_numpy(self):
        return self._df.to_numpy()

    def to_pandas(self):
        return self._df.copy()

    def __setitem__(self, key, value):
        if isinstance(key, slice) and value is dt.str64:
            self._df = self._df.astype(str).astype(object)
            return
        raise NotImplementedError(key)

    def __delitem__(self, key):
        rows, cols = key
        assert isinstance(rows, slice)
        self._df = self._df.drop(columns=list(cols))

    def __getitem__(self, key):
        rows, what = key
        assert isinstance(rows, slice)
        if isinstance(what, _Update):
            for name, e in what.exprs.items():
                assert isinstance(e, _IfElse) and isinstance(e.cond, _IsNA)
                col = self._df[e.cond.name]
                # datatable reads the all-integer rating columns as int32 with NA
                self._df[name] = col.fillna(e.a).astype(np.int32)
            return None
        raise NotImplementedError(what)


def _fread(path, header=None, columns=None):
    if columns is not None:
        # datatable semantic used by main.py:75 — the file's first line is a
        # header (one field shorter than the rows, generator.py:332) that is
        # dropped and replaced by `columns`.
        df = pd.read_csv(path, header=None, skiprows=1, names=list(columns))
    elif header is False:
        df = pd.read_csv(path, header=None)
        df.columns = ["C%d" % i for i in range(df.shape[1])]
    else:
        df = pd.read_csv(path)
    return Frame(df)


class _DT:
    str64 = _StrType()
    Frame = Frame
    fread = staticmethod(_fread)


dt = _DT()
str64 = dt.str64
fread = _fread
