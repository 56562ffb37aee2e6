"""`volumetricinterp [--validate] config_file` — same command line as the reference
(reference volumetricinterp/run_volumetricinterp.py:14-35)."""
import argparse

from .interpolate import Interpolate


def main(argv=None):
    parser = argparse.ArgumentParser(description='Fit a 3D analytic model to AMISR data (B200-native hot path).')
    parser.add_argument('config_file', help='configuration file, same keys as the reference example_config.ini')
    parser.add_argument('--validate', action='store_true',
                        help='fit only the [VALIDATE] STARTTIME..ENDTIME window (the reference then draws a PNG; '
                             'plotting is out of scope here: matplotlib/cartopy are not part of the hot path)')
    args = parser.parse_args(argv)
    interp = Interpolate(args.config_file)
    if args.validate:
        import configparser
        import datetime as dt
        cfg = configparser.ConfigParser()
        with open(args.config_file) as f:
            cfg.read_file(f)
        t0 = dt.datetime.strptime(cfg.get('VALIDATE', 'STARTTIME'), '%Y-%m-%dT%H:%M:%S')
        t1 = dt.datetime.strptime(cfg.get('VALIDATE', 'ENDTIME'), '%Y-%m-%dT%H:%M:%S')
        interp.calc_coeffs(starttime=t0, endtime=t1)
    else:
        interp.calc_coeffs()
    interp.saveh5()
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
