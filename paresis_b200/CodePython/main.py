#!/usr/bin/env python3
"""Driver for a speckle-based imaging simulation -- same flow and output tree as PARESIS's main.py
(main.py:20-115), runnable from this directory (or any workspace made by
``paresis_b200.workspace.make_workspace``):

    python main.py                       # the bundled example: Fil_Nylon_ID17, RayT, 1 position
    python main.py --experiment B200_2048_mono --points 20 --results /tmp/out

PARESIS's own main.py also runs unchanged against these modules; this one only adds flags.
"""
import argparse
import datetime
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from InputOutput.pagailleIO import save_image  # noqa: E402
from Experiment import Experiment  # noqa: E402


def run(args):
    time0 = time.time()
    exp_dict = {
        'experimentName': args.experiment,
        'filepath': os.path.join(args.results, args.experiment, ''),
        'overSampling': args.oversampling,       # integer; >= 2 for ray tracing, more for Fresnel
        'nbExpPoints': args.points,              # membrane positions, i.e. (Ir, Is) pairs
        'simulation_type': args.model,           # "RayT" | "Fresnel"
        'expID': datetime.datetime.now().strftime("%Y%m%d-%H%M%S"),
        'resultDtype': "float32",                # images are saved as float32 (pagailleIO): fetch them that way
    }
    if args.seed is not None:
        exp_dict['seed'] = args.seed
    os.makedirs(exp_dict['filepath'], exist_ok=True)
    ext = args.format

    print("\n\nINITIALIZING EXPERIMENT PARAMETERS AND GEOMETRIES\n*************************")
    experiment = Experiment(exp_dict)
    print("\nImages calculation\n*************************")
    root = bins = None
    for point in range(exp_dict['nbExpPoints']):
        experiment.myMembrane.myGeometry = []
        experiment.myMembrane.getMyGeometry(experiment.exp_dict['studyDimensions'], experiment.myMembrane.membranePixelSize,
                                            experiment.exp_dict['overSampling'], point, exp_dict['nbExpPoints'])
        print("\nCalculations point", point)
        if args.model == "Fresnel":
            sample, ref, propag, white = experiment.computeSampleAndReferenceImages_Fresnel(point)
            dark = None
        elif args.model == "RayT":
            sample, ref, propag, white, _dx, _dy, dark = experiment.computeSampleAndReferenceImages_RT(point)
        else:
            raise Exception("simulation Type not defined: ", args.model)
        if point == 0:
            root = exp_dict['filepath'] + ('Fresnel_' if args.model == "Fresnel" else 'RayTracing_') + exp_dict['expID'] + '/'
            os.mkdir(root)
            os.mkdir(root + 'membraneThickness/')
            edges = [experiment.mySource.mySpectrum[0][0]] + list(experiment.myDetector.det_param['myBinsThersholds'])
            if len(edges) == 2:
                bins = [root]                     # one energy bin: no per-bin sub-folders
            else:
                bins = ['%s%2.2d_%2.2dkev/' % (root, edges[b], edges[b + 1]) for b in range(len(sample))]
                for b in bins:
                    os.mkdir(b)
            for b in bins:
                for sub in ('ref/', 'sample/', 'propag/'):
                    os.mkdir(b + sub)
        tag = '%2.2d' % point
        save_image(experiment.myMembrane.myGeometry[0],
                   root + 'membraneThickness/' + args.experiment + '_sampling' + str(args.oversampling) + '_' + str(point) + ext)
        if dark is not None:
            save_image(dark, root + "DF" + ext)
        for b, folder in enumerate(bins):
            save_image(sample[b], folder + 'sample/sampleImage_' + exp_dict['expID'] + '_' + tag + ext)
            save_image(ref[b], folder + 'ref/ReferenceImage_' + exp_dict['expID'] + '_' + tag + ext)
            if point == 0:
                save_image(propag[b], folder + 'propag/PropagImage_' + exp_dict['expID'] + '_' + ext)
                save_image(white[b], folder + 'White_' + exp_dict['expID'] + '_' + ext)
    experiment.saveAllParameters(time0, exp_dict)
    print("\nfini")
    return root


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--experiment", default="Fil_Nylon_ID17")
    ap.add_argument("--results", default="../Results")
    ap.add_argument("--oversampling", type=int, default=2)
    ap.add_argument("--points", type=int, default=1)
    ap.add_argument("--model", choices=("RayT", "Fresnel"), default="RayT")
    ap.add_argument("--format", choices=(".tif", ".edf"), default=".tif")
    ap.add_argument("--seed", type=int, default=None)
    run(ap.parse_args())
