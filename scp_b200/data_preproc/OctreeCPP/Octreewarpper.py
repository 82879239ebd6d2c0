"""Drop-in for data_preproc/OctreeCPP/Octreewarpper.py: the same ctypes surface (``gen_octree``, ``COctree`` with
``len()``, ``[level]``, ``.node[i] -> Node{nodeid,octant,parent,oct,pos[3]}``, ``.code``), bound to the legacy C
symbols that libscp_b200.so exports in place of the prebuilt ``Octree_python_lib.so`` (Octreewarpper.py:17-39)."""
from ctypes import POINTER, c_double, c_int

import numpy as np

from ... import _lib

Node = _lib.LegacyNode
c_double_p = POINTER(c_double)


def _lib_handle():
    return _lib.require_device()


class OctCode:
    def __init__(self, adr):
        self.nodeAdr = adr
        self.Len = _lib_handle().int_size(adr)

    def __getitem__(self, i):
        L = self.Len
        if i >= L or i < -L:
            raise IndexError('Vector index out of range')
        return _lib_handle().int_get(self.nodeAdr, i + L if i < 0 else i)

    def __len__(self):
        return self.Len


class _Nodes:
    def __init__(self, adr):
        self.nodeAdr = adr
        self.Len = _lib_handle().Nodes_size(adr)

    def __getitem__(self, i):
        L = self.Len
        if i >= L or i < -L:
            raise IndexError('Vector index out of range')
        return _lib_handle().Nodes_get(self.nodeAdr, i + L if i < 0 else i).contents

    def __len__(self):
        return self.Len


class Level:
    def __init__(self, adr, i):
        self.Adr = adr
        self.node = _Nodes(adr)
        self.level = i + 1
        self.Len = len(self.node)

    def __getitem__(self, i):
        return self.node[i]

    def __len__(self):
        return self.Len


class COctree(object):
    def __init__(self):
        self.lib = _lib_handle()
        self.vector = self.lib.new_vector()
        self.code = None

    def __del__(self):
        try:
            self.lib.delete_vector(self.vector)
        except Exception:
            pass

    def __len__(self):
        return self.lib.vector_size(self.vector)

    def __getitem__(self, i):
        L = self.__len__()
        if i >= L or i < -L:
            raise IndexError('Vector index out of range')
        if i < 0:
            i += L
        return Level(self.lib.vector_get(self.vector, c_int(i)), i)

    def push(self, i):
        self.lib.vector_push_back(self.vector, c_int(i))

    def genOctree(self, p):
        data = np.ascontiguousarray(p).astype(np.double)
        adr = self.lib.genOctreeInterface(self.vector, data.ctypes.data_as(c_double_p), data.shape[0])
        if not adr:
            raise _lib.ScpError("genOctreeInterface: " + self.lib.scp_last_error().decode(errors="replace"))
        self.code = OctCode(adr)


def gen_octree(points):
    octree = COctree()
    octree.genOctree(points)
    return octree
