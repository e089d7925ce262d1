"""Mesh -> node-set pipeline: the host-side mirror of src/processmesh.jl and its helpers.

The reference reads CGNS files through HDF5.jl (third-party).  There is no HDF5 library in this image, so `Hdf5File`
below is a minimal read-only parser of the subset of the HDF5 file format those files use (superblock v2/v3, version-2
object headers with continuation blocks, compact link messages and dense link storage in a fractal heap, dataspace /
datatype / contiguous + compact layout messages, little-endian integer and IEEE float types).  `processmesh` follows
src/processmesh.jl:1-192 line by line; the two places where the reference calls the approximate HNSW search
(normal orientation :125-139, ghost offset :146-171) use the exact nearest neighbour instead -- on the GPU through
`rbffd_knn_device` when a context is given, else a numpy brute force (host preprocessing of O(boundary) queries).
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5File:
    """h5 = Hdf5File(path); h5["Base/dom-1/GridCoordinates/CoordinateX/ data"] -> numpy array (C order of the file)."""

    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        ver = b[8]
        if ver not in (2, 3):
            raise ValueError(f"HDF5 superblock version {ver} not supported (CGNS files written by HDF5 >= 1.8 use 2 or 3)")
        self.so, self.sl = b[9], b[10]
        if self.so != 8 or self.sl != 8:
            raise ValueError("only 8-byte offsets/lengths are supported")
        self.base = self._u(12, 8)
        self.root = self._u(36, 8)
        self._links = {}

    def _u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    # ---- object headers (version 2) ---------------------------------------------------------------------------
    def _messages(self, addr):
        """-> list of (type, bytes) of the object header at `addr`, continuation blocks followed."""
        b = self.b
        addr += self.base
        if b[addr:addr + 4] != b"OHDR" or b[addr + 4] != 2:
            raise ValueError("only version-2 object headers are supported")
        flags = b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szlen = 1 << (flags & 3)
        chunk = self._u(p, szlen)
        p += szlen
        blocks = [(p, p + chunk)]
        out = []
        while blocks:
            p, end = blocks.pop(0)
            while p + 4 <= end:
                mtype, msize, _mflags = b[p], self._u(p + 1, 2), b[p + 3]
                p += 4
                if flags & 0x04:
                    p += 2
                data = b[p:p + msize]
                p += msize
                if mtype == 0x10:                                     # continuation
                    coff, clen = struct.unpack("<QQ", data[:16])
                    coff += self.base
                    if b[coff:coff + 4] != b"OCHK":
                        raise ValueError("bad continuation block")
                    blocks.append((coff + 4, coff + clen - 4))        # signature ... checksum
                elif mtype != 0:
                    out.append((mtype, data))
        return out

    @staticmethod
    def _parse_link(data):
        """link message (type 6) -> (name, object header address or None for soft/external links, bytes consumed)"""
        ver, fl = data[0], data[1]
        if ver != 1:
            raise ValueError("link message version")
        p = 2
        ltype = 0
        if fl & 0x08:
            ltype = data[p]
            p += 1
        if fl & 0x04:
            p += 8
        if fl & 0x10:
            p += 1
        ln = 1 << (fl & 3)
        nlen = int.from_bytes(data[p:p + ln], "little")
        p += ln
        name = data[p:p + nlen].decode("utf-8", "replace")
        p += nlen
        if ltype == 0:
            return name, int.from_bytes(data[p:p + 8], "little"), p + 8
        if ltype == 1:                                                # soft link: length + string
            sl = int.from_bytes(data[p:p + 2], "little")
            return name, None, p + 2 + sl
        return name, None, len(data)

    def _dense_links(self, heap_addr):
        """links stored in a fractal heap (link info message): walk the managed direct blocks and parse the link
        messages that are packed back to back in them."""
        b = self.b
        h = heap_addr + self.base
        if b[h:h + 4] != b"FRHP":
            raise ValueError("fractal heap header expected")
        p = h + 5
        _idlen, filtlen = self._u(p, 2), self._u(p + 2, 2)
        p += 4
        hflags = b[p]
        p += 1
        p += 4                     # max managed object size
        p += 8 + 8                 # next huge id, huge b-tree address
        p += 8 + 8                 # free space, free-space manager address
        p += 8 * 4                 # managed space, allocated managed, iterator offset, number of managed objects
        p += 8 * 4                 # huge size/count, tiny size/count
        table_width = self._u(p, 2)
        p += 2
        start_block = self._u(p, 8)
        p += 8
        max_direct = self._u(p, 8)
        p += 8
        max_heap_bits = self._u(p, 2)
        p += 2
        p += 2                     # starting rows
        root = self._u(p, 8)
        p += 8
        cur_rows = self._u(p, 2)
        if filtlen:
            raise ValueError("filtered fractal heaps are not supported")
        off_bytes = (max_heap_bits + 7) // 8
        links = {}

        def direct(addr, size):
            a = addr + self.base
            if b[a:a + 4] != b"FHDB":
                raise ValueError("fractal heap direct block expected")
            q = a + 5 + 8 + off_bytes + (4 if hflags & 0x02 else 0)
            end = a + size
            while q + 4 < end and b[q] == 1:                          # link message version 1
                try:
                    name, oaddr, used = self._parse_link(b[q:end])
                except Exception:
                    break
                if oaddr is not None:
                    links[name] = oaddr
                q += used

        def indirect(addr, nrows):
            a = addr + self.base
            if b[a:a + 4] != b"FHIB":
                raise ValueError("fractal heap indirect block expected")
            q = a + 5 + 8 + off_bytes
            max_direct_rows = (max_direct // start_block).bit_length() + 1
            for r in range(nrows):
                size = start_block * (1 if r < 2 else 1 << (r - 1))
                for _ in range(table_width):
                    child = self._u(q, 8)
                    q += 8
                    if child == UNDEF:
                        continue
                    if r < max_direct_rows:
                        direct(child, size)
                    else:
                        raise ValueError("nested indirect fractal-heap blocks are not supported")

        if root != UNDEF:
            if cur_rows == 0:
                direct(root, start_block)
            else:
                indirect(root, cur_rows)
        return links

    def links(self, addr):
        if addr in self._links:
            return self._links[addr]
        out = {}
        for mtype, data in self._messages(addr):
            if mtype == 0x06:
                name, oaddr, _ = self._parse_link(data)
                if oaddr is not None:
                    out[name] = oaddr
            elif mtype == 0x02:                                       # link info: dense storage
                fl = data[1]
                p = 2 + (8 if fl & 1 else 0)
                heap = int.from_bytes(data[p:p + 8], "little")
                if heap != UNDEF:
                    out.update(self._dense_links(heap))
        self._links[addr] = out
        return out

    def resolve(self, path):
        addr = self.root
        for part in [s for s in path.split("/") if s != ""]:
            ln = self.links(addr)
            if part not in ln:
                raise KeyError(f"{part!r} not found under {path!r}; available: {sorted(ln)}")
            addr = ln[part]
        return addr

    def keys(self, path=""):
        return sorted(self.links(self.resolve(path)))

    # ---- datasets -----------------------------------------------------------------------------------------------
    def __getitem__(self, path):
        shape = dtype = layout = None
        for mtype, data in self._messages(self.resolve(path)):
            if mtype == 0x01:
                ver, rank = data[0], data[1]
                p = 8 if ver == 1 else 4
                shape = tuple(int.from_bytes(data[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))
            elif mtype == 0x03:
                cls, bits0, size = data[0] & 0x0F, data[1], int.from_bytes(data[4:8], "little")
                if bits0 & 1:
                    raise ValueError("big-endian datasets are not supported")
                if cls == 0:
                    dtype = np.dtype(("<i" if bits0 & 0x08 else "<u") + str(size))
                elif cls == 1:
                    dtype = np.dtype("<f" + str(size))
                elif cls == 3:
                    dtype = np.dtype("S" + str(size))
                else:
                    raise ValueError(f"datatype class {cls} not supported")
            elif mtype == 0x08:
                ver, cls = data[0], data[1]
                if ver not in (3, 4):
                    raise ValueError("data layout message version")
                if cls == 0:
                    n = int.from_bytes(data[2:4], "little")
                    layout = ("compact", data[4:4 + n])
                elif cls == 1:
                    layout = ("contiguous", int.from_bytes(data[2:10], "little"), int.from_bytes(data[10:18], "little"))
                else:
                    raise ValueError("chunked datasets are not supported")
        if shape is None or dtype is None or layout is None:
            raise KeyError(f"{path!r} is not a dataset")
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            raw = layout[1]
        else:
            if layout[1] == UNDEF:
                return np.zeros(shape, dtype)
            a = layout[1] + self.base
            raw = self.b[a:a + count * dtype.itemsize]
        return np.frombuffer(raw, dtype, count).reshape(shape).copy()


def _nearest(points, queries, ctx=None):
    """exact nearest neighbour (index into points, distance) of every query; GPU kNN when a context is given"""
    points = np.ascontiguousarray(points, np.float64)
    queries = np.ascontiguousarray(queries, np.float64)
    if ctx is not None:
        import torch
        dev = torch.device("cuda", ctx.device)
        P, Q = torch.from_numpy(points).to(dev), torch.from_numpy(queries).to(dev)
        idx = torch.empty((len(queries), 1), dtype=torch.int32, device=dev)
        d2 = torch.empty((len(queries), 1), dtype=torch.float64, device=dev)
        ctx.knn_device(P.data_ptr(), len(points), points.shape[1], 1, idx.data_ptr(), Q_ptr=Q.data_ptr(), NQ=len(queries),
                       d2_out_ptr=d2.data_ptr())
        ctx.synchronize()
        return idx.cpu().numpy()[:, 0].astype(np.int64), np.sqrt(d2.cpu().numpy()[:, 0])
    d2 = ((queries[:, None, :] - points[None, :, :]) ** 2).sum(-1)
    j = d2.argmin(1)
    return j, np.sqrt(d2[np.arange(len(queries)), j])


def processmesh(meshname, markernames, ctx=None):
    """src/processmesh.jl:1-192.  Returns (Y, y_point_mat, Y_idx_in, Y_idx_bc, Y_idx_bc_g, cells, bc_normals, bc_tangents):
    Y [M,2] = triangle centroids, boundary-edge midpoints per marker (CGNS element order), then the ghost nodes;
    index sets are 0-based Python ranges; cells = (triangles [nt,3], boundary edges [nb,2]) as 0-based vertex ids."""
    h5 = Hdf5File(meshname)
    zone = "Base/dom-1/"
    x = h5[zone + "GridCoordinates/CoordinateX/ data"].ravel()                  # extractcoordinates.jl:4-11
    y = h5[zone + "GridCoordinates/CoordinateY/ data"].ravel()
    P = np.stack([x, y], 1).astype(np.float64)
    tri = h5[zone + "TriElements/ElementConnectivity/ data"].ravel().astype(np.int64).reshape(-1, 3) - 1    # extractelements.jl:4-6
    centroids = P[tri].mean(1)                                                     # extractelements.jl:16
    int_range = h5[zone + "TriElements/ElementRange/ data"].ravel().astype(np.int64)
    Y_idx_in = range(int(int_range[0]) - 1, int(int_range[1]))                     # processmesh.jl:26-27
    bc_range = [h5[zone + m + "/ElementRange/ data"].ravel().astype(np.int64) for m in markernames]           # :55-60
    bc_max = int(max(r.max() for r in bc_range))
    Y = np.full((bc_max, 2), np.nan)
    Y[Y_idx_in.start:Y_idx_in.stop] = centroids                                    # :65
    Y_idx_bc, bc_lines, bc_normals, bc_tangents, bc_elems = [], [], [], [], []
    for m, r in zip(markernames, bc_range):                                        # :77-121
        Y_idx_bc.append(range(int(r[0]) - 1, int(r[1])))
        el = h5[zone + m + "/ElementConnectivity/ data"].ravel().astype(np.int64).reshape(-1, 2) - 1
        mid = P[el].mean(1)
        d = P[el[:, 1]] - P[el[:, 0]]                                              # calculatenormal.jl:7-11
        ln = np.hypot(-d[:, 1], d[:, 0])
        bc_normals.append(np.stack([-d[:, 1], d[:, 0]], 1) / ln[:, None])
        bc_tangents.append(d / ln[:, None])
        bc_lines.append(mid)
        bc_elems.append(el)
        Y[Y_idx_bc[-1].start:Y_idx_bc[-1].stop] = mid                              # :111
    # orient the normals out of the domain (:125-139); exact nearest interior node instead of HNSW
    first = Y[Y_idx_bc[0].start]
    j, _ = _nearest(Y[Y_idx_in.start:Y_idx_in.stop], first[None, :], ctx)
    orient = np.sign(np.dot(first - Y[Y_idx_in.start + int(j[0])], bc_normals[0][0]))
    bc_normals = [orient * v for v in bc_normals]
    bc_tangents = [orient * v for v in bc_tangents]
    # ghost offset = mean distance from the boundary nodes to their nearest interior node (:146-171)
    allbc = np.concatenate([Y[r.start:r.stop] for r in Y_idx_bc])
    _, dist = _nearest(Y[Y_idx_in.start:Y_idx_in.stop], allbc, ctx)
    offset = abs(dist.mean())
    shift = bc_max - int(int_range[1])                                             # :174
    Y_idx_bc_g = [range(r.start + shift, r.stop + shift) for r in Y_idx_bc]
    ghosts = np.full((bc_max - int(int_range[1]), 2), np.nan)
    for r, mid, nrm in zip(Y_idx_bc, bc_lines, bc_normals):                        # genghostnodes.jl:4, :175-181
        ghosts[r.start - int(int_range[1]):r.stop - int(int_range[1])] = mid + nrm * offset
    Y = np.concatenate([Y, ghosts])                                                # :186
    cells = (tri, np.concatenate(bc_elems))
    return Y, P.T.copy(), Y_idx_in, Y_idx_bc, Y_idx_bc_g, cells, bc_normals, bc_tangents
