"""Single-process stand-in for mpi4py.MPI: rank 0 of a world of size 1.

Only what /root/reference/Tools/Distributed_tools.py:10-11,79-91 touches.
Multi-rank behaviour of the reference is reproduced by oracle/ref_harness.py,
which calls the reference's own functions rank by rank and performs the
gather/sum/bcast of syn_cpus (Distributed_tools.py:77-92) in-process.
"""


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def bcast(self, obj, root=0):
        return obj

    def gather(self, obj, root=0):
        return [obj]

    def Gather(self, send, recv, root=0):
        import numpy as np
        recv[...] = np.asarray(send).reshape(recv.shape)

    def Gatherv(self, send, recv, root=0):
        import numpy as np
        recv[...] = np.asarray(send).reshape(recv.shape)

    def Barrier(self):
        pass


COMM_WORLD = _Comm()
