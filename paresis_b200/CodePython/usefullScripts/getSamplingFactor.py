"""Minimum oversampling for the Fresnel model -- drop-in for usefullScripts/getSamplingFactor.py
(criterion of Haggmark, Shaker & Hertz, IEEE TMI 40(2), 2020, as used at getSamplingFactor.py:17-26)."""
import numpy as np


def kevToLambda(energyInKev):
    return 1240. / (energyInKev * 1e3) * 1e-9


def is_overSampling_ok(exp_dict, pixel_size, energy):
    """Prints a warning when exp_dict['overSampling'] is below the Fresnel minimum; returns that minimum.
    (Upstream returns an unbound local for non-Fresnel runs; here that case returns None.)"""
    if exp_dict['simulation_type'] != "Fresnel":
        return None
    d1, d2, d3 = exp_dict['distSourceToMembrane'], exp_dict['distMembraneToObject'], exp_dict['distObjectToDetector']
    magnification = (d1 + d2 + d3) / (d1 + d2)
    min_dx = np.sqrt(kevToLambda(energy) * d3 / magnification) / 2
    min_oversampling = np.ceil(pixel_size / magnification / min_dx * 1e-6)
    if min_oversampling > exp_dict['overSampling']:
        print(f'/!\\/!\\ OVERSAMPLING FACTOR < MIN OVERSAMPLING FOR FRESNEL MODEL: {exp_dict["overSampling"]} < {min_oversampling}')
    return min_oversampling
