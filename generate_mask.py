"""Drop-in for the reference's generate_mask.py: `python generate_mask.py <pk|bk> <paramfile>` builds the smooth background
density n-bar(r) on the mesh from the randoms' n(z) and the angular mask and writes
`{output}nbar_{type}z%.3f_%.3f_g%d_%d_%d.npy` (generate_mask.py:77), normalised to the randoms (:140-142).
Same argv, .param keys, stdout milestones and early exit; the arithmetic lives in spectra_without_windows_b200/ingest.py."""
import configparser
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(argv):
    from spectra_without_windows_b200 import ingest
    from spectra_without_windows_b200.opt_utilities import box_center_of, load_randoms
    if len(argv) != 3:
        raise Exception("Need to specify spectrum type and parameter file!")
    spec_type, paramfile = str(argv[1]), str(argv[2])
    config = configparser.ConfigParser(interpolation=None, converters={"list": lambda x: [i.strip() for i in x.split(",")]})
    if not os.path.exists(paramfile):
        raise Exception("Parameter file does not exist!")
    config.read(paramfile)
    string = config["sample"]["type"]
    outdir = str(config["directories"]["output"])
    ZMIN, ZMAX = float(config["sample"]["z_min"]), float(config["sample"]["z_max"])
    h_fid, OmegaM_fid = float(config["parameters"]["h_fid"]), float(config["parameters"]["OmegaM_fid"])
    boxsize_grid = np.array(config.getlist("sample", "box"), dtype=float)
    if spec_type == "pk":
        grid_3d = np.array(config.getlist("pk-binning", "grid"), dtype=int)
    elif spec_type == "bk":
        grid_3d = np.array(config.getlist("bk-binning", "grid"), dtype=int)
    else:
        raise Exception("Spectrum type must be 'pk' or 'bk'!")
    maskfile = str(config["catalogs"]["mask_file"])
    if maskfile.strip().lower() not in ("", "none") and not os.path.exists(maskfile):
        raise Exception("Mask file does not exist!")
    if not os.path.exists(outdir):
        os.makedirs(outdir)
    print("\n###################### PARAMETERS ######################\n")
    print("Data-type: %s" % string)
    print("\nFiducial h = %.3f" % h_fid)
    print("Fiducial Omega_m = %.3f" % OmegaM_fid)
    print("Redshift = [%.3f, %.3f]" % (ZMIN, ZMAX))
    print("\nBoxsize = [%.3e, %.3e, %.3e] Mpc/h" % (boxsize_grid[0], boxsize_grid[1], boxsize_grid[2]))
    print("Grid = [%d, %d, %d]" % (grid_3d[0], grid_3d[1], grid_3d[2]))
    print("Output Directory: %s" % outdir)
    print("\n########################################################")
    init = time.time()
    outfile = ingest.nbar_file_name(outdir, string, ZMIN, ZMAX, grid_3d)
    if os.path.exists(outfile):
        print("Mask already completed; exiting!\n")
        sys.exit()
    cosmo_coord = ingest.Cosmology(h=h_fid).match(Omega0_m=OmegaM_fid)
    print("\n## Loading data and randoms")
    randoms = load_randoms(config, cosmo_coord, fkp_weights=False)
    if "Z" not in randoms:
        raise Exception("the randoms need a redshift column 'Z' for the n(z) histogram")
    # cell centres exactly as load_coord_grids builds them (src/opt_utilities.py:428-430)
    bc = box_center_of(randoms)
    from spectra_without_windows_b200 import engine
    H = boxsize_grid / grid_3d
    offset = bc + 0.5 * H
    r_axes = [(engine.signed_index(int(grid_3d[i])) * H[i]).astype("f8") + offset[i] for i in range(3)]
    print("\n## Creating RA, Dec, z co-ordinate grid")
    print("## Creating n(z) interpolator from random particles")
    print("## Defining angular mask")
    mask = None if maskfile.strip().lower() in ("", "none") else ingest.MangleMask(maskfile)
    print("## Computing total mask")
    nbar_mask = ingest.build_nbar_mask(r_axes, boxsize_grid, grid_3d, randoms["Z"].v, randoms["WEIGHT"].v, ZMIN, ZMAX,
                                       cosmo_coord, mask)
    np.save(outfile, nbar_mask)
    duration = time.time() - init
    print("Output map saved to %s. Exiting after %d seconds (%d minutes)\n" % (outfile, duration, duration // 60))


if __name__ == "__main__":
    main(sys.argv)
