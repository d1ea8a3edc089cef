"""Image input/output -- drop-in for InputOutput/pagailleIO.py (pagailleIO.py:13-134).

Same function names; the float32 TIFF / EDF containers are written by the in-repo writers
instead of fabio, so the files open in the same downstream tools.
"""
import os

import numpy as np

import _paresis_path  # noqa: F401
from paresis_b200.hostio import imageio


def openImage(filename):
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
    imageio.write_tiff(filename, np.asarray(image).astype(np.float32 if bit == 32 else np.uint16))


def saveEdf(data, filename):
    print(filename)
    imageio.write_edf(filename, np.asarray(data).astype(np.float32))


def save_image(data, filename):
    ext = filename.split(".")[-1]
    if ext in ("tif", "tiff"):
        save_tif_image(data, filename)
    if ext == "edf":
        saveEdf(data, filename)
