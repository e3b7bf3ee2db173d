#!/usr/bin/env python3
"""Command-line entry point, same contract as the reference's
code/berkeley_hydro_main.py:  --params P.json [--data D.csv].

The JSON must hold the twelve top-level keys the reference requires
(berkeley_hydro_main.py:40-43); the CSV has four header-less columns
ID, Datenum, Precipitation_cm, WTD_m (:115-116).  Any exception ends the
program with exit code 1, as in the reference.
"""
import json
import sys
from pathlib import Path

import pandas as pd

from hydromodel_b200.src.simulation import Simulation

REQUIRED_KEYS = ("Trees", "Well_No", "Output_Name", "IC_Filename", "Data_Filename",
                 "Water_Content", "Environmental", "Soil_Properties", "Site_Information",
                 "Simulation_Flags", "Hydrological_Model", "Hydraulic_Conductivity")


def validateInputParametersFile(filename):
    """Load the JSON and check that every required key is present."""
    with open(filename, "r") as fh:
        model_params = json.load(fh)
    for key in REQUIRED_KEYS:
        if key not in model_params:
            raise ValueError(f" Key: {key}, is not given.")
    print(" Model parameters are given correctly.")
    return model_params


def main(params_file=None, data_file=None):
    if params_file is None:
        print(" The simulation can't run without input parameters.")
        sys.exit(1)
    try:
        params = validateInputParametersFile(Path(params_file))
    except ValueError as err:
        print(err)
        sys.exit(1)
    data_file = Path(data_file) if data_file is not None else Path(params["Data_Filename"])
    print(f" Simulation water data file: {data_file}")
    try:
        with open(data_file, "r") as fh:
            water_data = pd.read_csv(fh, names=["ID", "Datenum", "Precipitation_cm", "WTD_m"])
        name = params["Output_Name"] if params["Output_Name"] is not None else "Sim_01"
        sim = Simulation(name)
        sim.setupModel(params, water_data)
        sim.run()
        sim.saveResults()
    except Exception as err:          # noqa: BLE001 - reference behaviour
        print(err)
        sys.exit(1)


if __name__ == "__main__":
    if len(sys.argv) <= 1:
        sys.exit("Error: Not enough input parameters.")
    import argparse
    parser = argparse.ArgumentParser(description=" Berkeley Hydrological Simulation ")
    parser.add_argument("--params", help=" Input file (.json) with simulation parameters.")
    parser.add_argument("--data", help=" Input file (.csv) with simulation data (e.g.: precipitation, wtd).")
    cli = parser.parse_args()
    main(cli.params, cli.data)
    print(" Simulation completed.")
