"""Sample thickness maps -- drop-in for Samples/createSampGeom.py.

Sphere and cylinder (the bundled configurations) are rendered by CUDA kernels; image-stack and
template geometries are host-side conveniences.  The two multi-sphere demo phantoms
(createSampGeom.py:110-260) are not part of the accelerated path.
"""
import glob

import numpy as np

import _paresis_path  # noqa: F401
from InputOutput.pagailleIO import openImage
from paresis_b200 import geometry
from paresis_b200.hostio import xmlparams

_XML = "xmlFiles/Samples.xml"


def _sample_entry(myName):
    entry = xmlparams.find_entry(_XML, "sample", myName)
    if entry is None:
        raise ValueError("Sample not found in the xml file")
    return entry


def CreateSampleSphere(myName, dimX, dimY, pixelSize):
    """Centred sphere (createSampGeom.py:15-53): [1, dimX, dimY] thickness in metres + report dict."""
    radius = _sample_entry(myName).get("myRadius", float)
    if radius / pixelSize * 2 > max(dimX, dimY):
        print("/!\\ Sphere size bigger than the field of view!")
    return geometry.sample_sphere(radius, dimX, dimY, pixelSize), {'Sphere_radius': (radius, 'um')}


def CreateSampleCylindre(myName, dimX, dimY, pixelSize):
    """Rotated cylinder (createSampGeom.py:56-107), same pixels as the imutils/OpenCV rotation."""
    entry = _sample_entry(myName)
    radius, orientation = entry.get("myRadius", float), entry.get("myOrientation", float)
    if radius / pixelSize * 2 > max(dimX, dimY):
        print("/!\\ Cylinder size bigger than the field of view!")
    geom = geometry.sample_cylinder(radius, orientation, dimX, dimY, pixelSize)
    return geom, {'Cylinder_radius': (radius, 'um'), 'Cylinder_orientation': (orientation, 'degree')}


def loadSampleGeometryFromImages(myGeometryFolder, dimX, dimY, pixsize):
    """One thickness image (metres) per material (createSampGeom.py:263-293)."""
    paths = sorted(glob.glob(myGeometryFolder + "/*.tif") + glob.glob(myGeometryFolder + "/*.tiff") +
                   glob.glob(myGeometryFolder + "/*.edf"))
    print(f'Your loaded geometry comprises thickness maps for {len(paths)} materials')
    if not paths:
        raise Exception("The sample geometry you are trying to load does not exist or is incorrectly named:", myGeometryFolder)
    return [openImage(p) for p in paths], {'myGeometryFolder': (myGeometryFolder, '')}


def CreateYourSampleGeometry(myName, dimX0, dimY0, pixelSize):
    """Editable template (createSampGeom.py:296-324): one uniform 5 um layer."""
    print(f'Creating your own geometry for sample {myName}')
    thickness = 5 * 1e-6
    return np.ones((1, dimX0, dimY0)) * thickness, {'geometry thickness': (thickness, 'um'),
                                                    'geometry other parameter': ("unitlessParameter", '')}


def _not_accelerated(name):
    def fn(*args, **kwargs):
        raise NotImplementedError("%s is a fixed demo phantom outside the accelerated hot path (SURVEY.md section 2, #9); "
                                  "render it once with PARESIS and load it with loadSampleGeometryFromImages" % name)
    fn.__name__ = name
    return fn


CreateSampleSpheresInCylinder = _not_accelerated("CreateSampleSpheresInCylinder")
CreateSampleSpheresInParallelepiped = _not_accelerated("CreateSampleSpheresInParallelepiped")


def getText(node):
    return xmlparams.text_of(node)
