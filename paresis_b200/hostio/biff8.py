"""Minimal read-only BIFF8 (.xls) workbook reader.

PARESIS keeps its delta/beta tables and tube spectra in legacy Excel files and reads
them through ``xlrd`` (reference ``Sample.py:108-146``, ``Detector.py:131-160``,
``Source.py:131-240``).  ``xlrd`` is not part of this image, so the host side carries
its own reader: OLE2 compound-document container -> ``Workbook`` stream -> the handful
of cell records those tables use (NUMBER, RK, MULRK, LABELSST, LABEL, FORMULA results).

The object model mirrors the slice of the xlrd API the reference touches:
``open_workbook(path).sheets()[k].cell(r, c).value`` plus ``.nrows`` / ``.ncols``.
Empty cells read as ``''`` exactly like xlrd's XL_CELL_EMPTY.
"""
import struct

_OLE_MAGIC = bytes.fromhex("D0CF11E0A1B11AE1")
_END = 0xFFFFFFFA  # any FAT entry >= this terminates a chain


class Cell:
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value


class Sheet:
    def __init__(self, name, cells):
        self.name = name
        self._cells = cells
        self.nrows = 1 + max((r for r, _ in cells), default=-1)
        self.ncols = 1 + max((c for _, c in cells), default=-1)

    def cell(self, row, col):
        if row < 0 or col < 0 or row >= self.nrows or col >= self.ncols:
            raise IndexError("cell (%d, %d) outside sheet %r" % (row, col, self.name))
        return Cell(self._cells.get((row, col), ""))

    def cell_value(self, row, col):
        return self.cell(row, col).value


class Workbook:
    def __init__(self, sheets):
        self._sheets = sheets

    def sheets(self):
        return list(self._sheets)

    def sheet_by_index(self, k):
        return self._sheets[k]


def _workbook_stream(blob):
    if blob[:8] != _OLE_MAGIC:
        raise ValueError("not an OLE2 compound document")
    sector = 1 << struct.unpack_from("<H", blob, 30)[0]
    mini = 1 << struct.unpack_from("<H", blob, 32)[0]
    (n_fat, dir_start, _sig, mini_cutoff, minifat_start, _n_minifat,
     difat_start, n_difat) = struct.unpack_from("<8I", blob, 44)

    def sect(i):
        off = (i + 1) * sector
        return blob[off:off + sector]

    per = sector // 4
    difat = list(struct.unpack_from("<109I", blob, 76))
    nxt = difat_start
    for _ in range(n_difat):
        if nxt >= _END:
            break
        entries = struct.unpack("<%dI" % per, sect(nxt))
        difat.extend(entries[:-1])
        nxt = entries[-1]
    fat = []
    for s in difat[:n_fat]:
        fat.extend(struct.unpack("<%dI" % per, sect(s)))

    def chain(table, start):
        out, seen = [], 0
        while start < _END:
            out.append(start)
            start = table[start]
            seen += 1
            if seen > len(table):
                raise ValueError("cyclic sector chain")
        return out

    directory = b"".join(sect(s) for s in chain(fat, dir_start))
    root_start = struct.unpack_from("<I", directory, 116)[0]
    for off in range(0, len(directory), 128):
        ent = directory[off:off + 128]
        nlen = struct.unpack_from("<H", ent, 64)[0]
        if nlen < 2:
            continue
        name = ent[:nlen - 2].decode("utf-16le")
        if name not in ("Workbook", "Book"):
            continue
        start, size = struct.unpack_from("<II", ent, 116)
        if size >= mini_cutoff:
            return b"".join(sect(s) for s in chain(fat, start))[:size]
        ministream = b"".join(sect(s) for s in chain(fat, root_start))
        minifat = []
        for s in chain(fat, minifat_start):
            minifat.extend(struct.unpack("<%dI" % per, sect(s)))
        return b"".join(ministream[s * mini:(s + 1) * mini] for s in chain(minifat, start))[:size]
    raise ValueError("no Workbook stream in compound document")


def _rk(v):
    if v & 2:
        i = v >> 2
        if v & 0x80000000:
            i -= 1 << 30
        x = float(i)
    else:
        x = struct.unpack("<d", struct.pack("<II", 0, v & 0xFFFFFFFC))[0]
    return x / 100.0 if v & 1 else x


def _shared_strings(chunks):
    """Decode the SST record; a string may straddle CONTINUE records (fresh flag byte)."""
    _total, unique = struct.unpack_from("<II", chunks[0], 0)
    ci, p = 0, 8
    out = []
    for _ in range(unique):
        if p >= len(chunks[ci]):
            ci, p = ci + 1, 0
        cch = struct.unpack_from("<H", chunks[ci], p)[0]
        flags = chunks[ci][p + 2]
        p += 3
        runs = ext = 0
        if flags & 8:
            runs = struct.unpack_from("<H", chunks[ci], p)[0]
            p += 2
        if flags & 4:
            ext = struct.unpack_from("<I", chunks[ci], p)[0]
            p += 4
        wide = flags & 1
        text = []
        while cch:
            if p >= len(chunks[ci]):
                ci += 1
                wide = chunks[ci][0] & 1
                p = 1
            room = len(chunks[ci]) - p
            if wide:
                n = min(cch, room // 2)
                text.append(chunks[ci][p:p + 2 * n].decode("utf-16le"))
                p += 2 * n
            else:
                n = min(cch, room)
                text.append(chunks[ci][p:p + n].decode("latin-1"))
                p += n
            cch -= n
        skip = 4 * runs + ext
        while skip:
            if p >= len(chunks[ci]):
                ci, p = ci + 1, 0
            n = min(skip, len(chunks[ci]) - p)
            p += n
            skip -= n
        out.append("".join(text))
    return out


def _parse(stream):
    records = []
    pos = 0
    while pos + 4 <= len(stream):
        op, ln = struct.unpack_from("<HH", stream, pos)
        records.append((op, stream[pos + 4:pos + 4 + ln]))
        pos += 4 + ln

    sst, names, sheets = [], [], []
    cur = None
    i = 0
    while i < len(records):
        op, data = records[i]
        if op == 0x00FC:  # SST (+ CONTINUE)
            chunks = [data]
            i += 1
            while i < len(records) and records[i][0] == 0x003C:
                chunks.append(records[i][1])
                i += 1
            sst = _shared_strings(chunks)
            continue
        if op == 0x0085:  # BOUNDSHEET
            cch, fl = data[6], data[7]
            names.append(data[8:8 + cch * (2 if fl & 1 else 1)].decode("utf-16le" if fl & 1 else "latin-1"))
        elif op == 0x0809:  # BOF
            kind = struct.unpack_from("<H", data, 2)[0]
            if kind == 0x0010:
                cur = {}
                sheets.append(cur)
            elif kind != 0x0005:
                cur = None
        elif cur is not None:
            if op == 0x0203:  # NUMBER
                r, c = struct.unpack_from("<HH", data, 0)
                cur[(r, c)] = struct.unpack_from("<d", data, 6)[0]
            elif op == 0x027E:  # RK
                r, c = struct.unpack_from("<HH", data, 0)
                cur[(r, c)] = _rk(struct.unpack_from("<I", data, 6)[0])
            elif op == 0x00BD:  # MULRK
                r, c0 = struct.unpack_from("<HH", data, 0)
                for k in range((len(data) - 6) // 6):
                    cur[(r, c0 + k)] = _rk(struct.unpack_from("<I", data, 6 + 6 * k)[0])
            elif op == 0x00FD:  # LABELSST
                r, c, _xf, idx = struct.unpack_from("<HHHI", data, 0)
                cur[(r, c)] = sst[idx]
            elif op == 0x0204:  # LABEL
                r, c = struct.unpack_from("<HH", data, 0)
                cch, fl = struct.unpack_from("<H", data, 6)[0], data[8]
                cur[(r, c)] = data[9:9 + cch * (2 if fl & 1 else 1)].decode("utf-16le" if fl & 1 else "latin-1")
            elif op == 0x0006:  # FORMULA: keep cached numeric results
                r, c = struct.unpack_from("<HH", data, 0)
                if data[12:14] != b"\xff\xff":
                    cur[(r, c)] = struct.unpack_from("<d", data, 6)[0]
        i += 1
    names += ["Sheet%d" % (k + 1) for k in range(len(names), len(sheets))]
    return Workbook([Sheet(n, cells) for n, cells in zip(names, sheets)])


def open_workbook(path):
    with open(path, "rb") as fh:
        return _parse(_workbook_stream(fh.read()))
