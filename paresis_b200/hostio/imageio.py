"""Float32 TIFF and EDF read/write without fabio.

PARESIS stores every result as a float32 image through fabio (InputOutput/pagailleIO.py:100-134).
These are the two containers it uses, written from scratch: baseline TIFF 6.0 (little endian, one
uncompressed strip, SampleFormat = IEEE float) and ESRF EDF (ASCII header padded to 512-byte
blocks + raw data).
"""
import struct

import numpy as np

_TIFF_TYPES = {np.dtype("float32"): (32, 3), np.dtype("uint16"): (16, 1), np.dtype("uint8"): (8, 1),
               np.dtype("int32"): (32, 2), np.dtype("float64"): (64, 3)}


def write_tiff(path, image):
    img = np.ascontiguousarray(image)
    if img.ndim != 2 or img.dtype not in _TIFF_TYPES:
        raise ValueError("write_tiff: need a 2-D array of float32/float64/uint16/uint8/int32")
    bits, fmt = _TIFF_TYPES[img.dtype]
    h, w = img.shape
    data = img.astype(img.dtype.newbyteorder("<")).tobytes()
    tags = [
        (256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, bits), (259, 3, 1, 1), (262, 3, 1, 1),
        (273, 4, 1, 8), (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, len(data)), (339, 3, 1, fmt),
    ]
    ifd_off = 8 + len(data) + (len(data) & 1)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<2sHI", b"II", 42, ifd_off))
        fh.write(data)
        if len(data) & 1:
            fh.write(b"\0")
        fh.write(struct.pack("<H", len(tags)))
        for tag, typ, cnt, val in tags:
            fh.write(struct.pack("<HHI", tag, typ, cnt) + (struct.pack("<HH", val, 0) if typ == 3 else struct.pack("<I", val)))
        fh.write(struct.pack("<I", 0))


def read_tiff(path):
    with open(path, "rb") as fh:
        blob = fh.read()
    bo = {b"II": "<", b"MM": ">"}[blob[:2]]
    magic, ifd = struct.unpack(bo + "HI", blob[2:8])
    if magic != 42:
        raise ValueError("not a TIFF file: " + path)
    n = struct.unpack_from(bo + "H", blob, ifd)[0]
    tags = {}
    for k in range(n):
        tag, typ, cnt, raw = struct.unpack_from(bo + "HHI4s", blob, ifd + 2 + 12 * k)
        size = {1: 1, 3: 2, 4: 4}.get(typ, 4) * cnt
        src = raw if size <= 4 else blob[struct.unpack(bo + "I", raw)[0]:][:size]
        code = {1: "B", 3: "H", 4: "I"}.get(typ, "I")
        tags[tag] = struct.unpack_from(bo + code * cnt, src)
    if tags.get(259, (1,))[0] != 1:
        raise ValueError("compressed TIFF not supported")
    w, h, bits, fmt = tags[256][0], tags[257][0], tags[258][0], tags.get(339, (1,))[0]
    dtype = {(32, 3): "f4", (64, 3): "f8", (16, 1): "u2", (8, 1): "u1", (32, 2): "i4", (32, 1): "u4", (16, 2): "i2"}[(bits, fmt)]
    data = b"".join(blob[o:o + c] for o, c in zip(tags[273], tags[279]))
    return np.frombuffer(data, dtype=np.dtype(dtype).newbyteorder(bo), count=w * h).reshape(h, w).astype(dtype)


_EDF_TYPES = {np.dtype("float32"): "FloatValue", np.dtype("float64"): "DoubleValue", np.dtype("uint16"): "UnsignedShort",
              np.dtype("int32"): "SignedInteger", np.dtype("uint8"): "UnsignedByte"}


def write_edf(path, image):
    img = np.ascontiguousarray(image)
    if img.ndim != 2 or img.dtype not in _EDF_TYPES:
        raise ValueError("write_edf: unsupported array")
    h, w = img.shape
    fields = [("HeaderID", "EH:000001:000000:000000"), ("Image", "1"), ("ByteOrder", "LowByteFirst"),
              ("DataType", _EDF_TYPES[img.dtype]), ("Dim_1", str(w)), ("Dim_2", str(h)), ("Size", str(img.nbytes))]
    head = "{\n" + "".join("%s = %s ;\n" % kv for kv in fields)
    pad = -(len(head) + 2) % 512
    head = head + " " * pad + "}\n"
    with open(path, "wb") as fh:
        fh.write(head.encode("ascii"))
        fh.write(img.astype(img.dtype.newbyteorder("<")).tobytes())


def read_edf(path):
    with open(path, "rb") as fh:
        blob = fh.read()
    end = blob.index(b"}\n") + 2
    fields = {}
    for line in blob[:end].decode("ascii", "replace").split(";"):
        if "=" in line:
            k, v = line.split("=", 1)
            fields[k.strip(" {\n\r")] = v.strip()
    dtype = {v: k for k, v in _EDF_TYPES.items()}[fields["DataType"]]
    bo = "<" if fields.get("ByteOrder", "LowByteFirst") == "LowByteFirst" else ">"
    w, h = int(fields["Dim_1"]), int(fields["Dim_2"])
    return np.frombuffer(blob, dtype=dtype.newbyteorder(bo), count=w * h, offset=end).reshape(h, w).astype(dtype)


def open_image(path):
    path = str(path)
    return read_edf(path) if path.lower().endswith(".edf") else read_tiff(path)


# ---------------------------------------------------------------------------- asynchronous writer
class AsyncWriter:
    """One background thread that encodes and writes images, so that disk I/O overlaps the GPU work of the next
    membrane position (PARESIS writes ~6 images per position synchronously: main.py:98-110).  ``submit`` takes ownership
    of a private float32 / uint16 copy made on the caller's thread; at most ``depth`` images wait in memory.  Errors
    surface on the next ``submit`` / ``flush`` (and at interpreter exit)."""

    def __init__(self, depth=16):
        import queue
        import threading
        self._q = queue.Queue(maxsize=depth)
        self._error = None
        self._thread = threading.Thread(target=self._run, name="paresis-image-writer", daemon=True)
        self._thread.start()

    def _run(self):
        while True:
            item = self._q.get()
            try:
                if item is None:
                    return
                fn, path, image = item
                if self._error is None:
                    fn(path, image)
            except BaseException as exc:          # kept for the submitting thread
                self._error = exc
            finally:
                self._q.task_done()

    def _raise(self):
        if self._error is not None:
            exc, self._error = self._error, None
            raise exc

    def submit(self, fn, path, image):
        self._raise()
        self._q.put((fn, path, image))

    def flush(self):
        """Block until everything submitted so far is on disk."""
        self._q.join()
        self._raise()


_writer = None


def writer():
    """The process-wide writer thread (created on first use, flushed at interpreter exit)."""
    global _writer
    if _writer is None:
        import atexit
        _writer = AsyncWriter()
        atexit.register(_writer.flush)
    return _writer


def flush_writes():
    if _writer is not None:
        _writer.flush()
