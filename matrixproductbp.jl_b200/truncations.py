"""SVD truncation policies (TensorTrains.jl names re-exported at /root/reference/src/MatrixProductBP.jl:42,69)."""


class SVDTrunc:
    kind = 0
    d = 0
    eps = 0.0


class TruncBond(SVDTrunc):
    def __init__(self, mprime):
        self.kind, self.d, self.eps = 0, int(mprime), 0.0

    def __repr__(self):
        return f"TruncBond({self.d})"


class TruncBondMax(TruncBond):
    def __repr__(self):
        return f"TruncBondMax({self.d})"


class TruncThresh(SVDTrunc):
    def __init__(self, eps):
        self.kind, self.d, self.eps = 1, 0, float(eps)

    def __repr__(self):
        return f"TruncThresh({self.eps})"


class TruncBondThresh(SVDTrunc):
    def __init__(self, mprime, eps=0.0):
        self.kind, self.d, self.eps = 2, int(mprime), float(eps)

    def __repr__(self):
        return f"TruncBondThresh({self.d},{self.eps})"
