"""Sample thickness maps -- drop-in for Samples/createSampGeom.py.

Sphere, cylinder and the two multi-sphere phantoms are rendered by CUDA kernels; image-stack and
template geometries are host-side conveniences.
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


def CreateSampleSpheresInCylinder(myName, dimX, dimY, pixelSize):
    """Two 500 um spheres in a vertical cylinder, three materials (createSampGeom.py:110-172)."""
    r0 = 500
    geom = geometry.sample_two_spheres(0, dimX, dimY, pixelSize)
    pos1, pos2 = int(np.round(r0 * 3 / pixelSize)), int(np.round(r0 * 7 / pixelSize))
    return geom, {'Spheres_radius': (r0, 'um'), 'Cylinder_radius': (r0 * 2, 'um'),
                  'Position_Sphere_1': (pos1 * pixelSize, 'um'), 'Position_Sphere_2': (pos2 * pixelSize, 'um')}


def CreateSampleSpheresInParallelepiped(myName, dimX0, dimY0, pixelSize):
    """Two 500 um spheres in a tilted, rounded parallelepiped, three materials (createSampGeom.py:174-260)."""
    r0 = 500
    margin = max(dimX0, dimY0) // 2
    dim_x = dimX0 + 2 * margin
    if abs(dim_x * 3 // 5 - dim_x * 2 // 5) < r0 / pixelSize * 2:
        print("/!\\ sample spheres overlapping!")
    geom = geometry.sample_two_spheres(1, dimX0, dimY0, pixelSize)
    return geom, {'Spheres_radius': (r0, 'um'), 'Parallelepipede_size': (r0 * 2, 'um'),
                  'Position_Sphere_1': (dim_x * 2 // 5 * pixelSize, 'um'), 'Position_Sphere_2': (dim_x * 3 // 5 * pixelSize, 'um')}


def getText(node):
    return xmlparams.text_of(node)
