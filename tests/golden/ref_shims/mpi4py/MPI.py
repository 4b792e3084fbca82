class _Comm:
    rank = 0
    size = 1
    def Get_rank(self): return 0
    def Get_size(self): return 1
    def bcast(self, x, root=0): return x
    def Barrier(self): pass
COMM_WORLD = _Comm()
