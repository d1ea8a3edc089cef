"""Image input/output -- drop-in for InputOutput/pagailleIO.py (pagailleIO.py:13-134).

Same function names; the float32 TIFF / EDF containers are written by the in-repo writers
instead of fabio, so the files open in the same downstream tools.

Writes are ASYNCHRONOUS: ``save_image`` converts the image to its file dtype (a private copy), hands it to one
background thread (``paresis_b200.hostio.imageio.AsyncWriter``) and returns, so the disk I/O of a membrane position
overlaps the GPU work of the next one -- also under PARESIS's unmodified main.py, which only calls ``save_image``.
Files are complete after ``wait_for_writes()`` and, at the latest, at interpreter exit; ``openImage`` waits first.
``PARESIS_B200_SYNC_IO=1`` restores synchronous writes.
"""
import os

import numpy as np

import _paresis_path  # noqa: F401
from paresis_b200.hostio import imageio


_SYNC = os.environ.get("PARESIS_B200_SYNC_IO", "") == "1"


def wait_for_writes():
    """Block until every image handed to save_image / save_tif_image / saveEdf so far is on disk."""
    imageio.flush_writes()


def _write(fn, filename, image):
    if _SYNC:
        fn(filename, image)
    else:
        imageio.writer().submit(fn, filename, image)


def openImage(filename):
    wait_for_writes()
    return imageio.open_image(filename)


def openSeq(filenames):
    if len(filenames) == 0:
        raise Exception('spytlabIOError')
    first = openImage(filenames[0])
    stack = np.zeros((len(filenames),) + first.shape, dtype=np.float32)
    for i, name in enumerate(filenames):
        stack[i] = openImage(name)
    return stack


def remove_filename_in_path(path):
    splitter = "\\" if len(path.split("\\")) > 1 else "/"
    return "".join(part + splitter for part in path.split(splitter)[:-1])


def create_directory(path):
    if path and not os.path.exists(path):
        os.makedirs(path)


def save_tif_image(image, filename, bit=32, header=None):
    """float32 (bit=32) or uint16 TIFF (pagailleIO.py:100-123); `header` is accepted and ignored."""
    create_directory(remove_filename_in_path(filename))
    _write(imageio.write_tiff, filename, np.asarray(image).astype(np.float32 if bit == 32 else np.uint16))


def saveEdf(data, filename):
    print(filename)
    _write(imageio.write_edf, filename, np.asarray(data).astype(np.float32))


def save_image(data, filename):
    ext = filename.split(".")[-1]
    if ext in ("tif", "tiff"):
        save_tif_image(data, filename)
    if ext == "edf":
        saveEdf(data, filename)
