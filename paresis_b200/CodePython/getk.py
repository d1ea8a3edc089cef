"""Wavenumber from photon energy -- drop-in for the reference's getk.py (getk.py:12-20)."""
import _paresis_path  # noqa: F401
from paresis_b200.hostmath import wavenumber


def getk(energy):
    """energy in eV -> k in 1/m, with the reference's rounded constants
    (h = 6.626e-34, c = 2.998e8, e = 1.6e-19)."""
    return wavenumber(energy)


if __name__ == "__main__":
    print("k=", getk(25000))
