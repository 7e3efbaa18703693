"""Dataset CSV writer (mirror of lens/tools/create_data_csv.py:6-58)."""
import csv
import os

import numpy as np


def haversine(lon1, lat1, lon2, lat2):
    """Great-circle distance in metres (create_data_csv.py:6-18)."""
    earth_radius_km = 6371.0
    lon1, lat1, lon2, lat2 = map(np.radians, [lon1, lat1, lon2, lat2])
    a = np.sin((lat2 - lat1) / 2) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) / 2) ** 2
    return earth_radius_km * (2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a))) * 1000


def create_csv_from_images(folder_path, csv_file_path, gps_path=None, fps=60, distance_threshold=100):
    """One row per *.png of folder_path in sorted order: Image_name, index[, gps_coordinate]
    (create_data_csv.py:20-58).  The GPS column needs the NMEA reader (pynmea2), as in the reference."""
    png_files = sorted(f for f in os.listdir(folder_path) if f.endswith(".png"))
    gps = None
    if gps_path is not None:
        from .read_gps import get_gps
        gps = get_gps(gps_path)
    with open(csv_file_path, "w", newline="") as fh:
        writer = csv.writer(fh)
        if gps is None:
            writer.writerow(["Image_name", "index"])
            for index, image_name in enumerate(png_files):
                writer.writerow([image_name, index])
            return
        writer.writerow(["Image_name", "index", "gps_coordinate"])
        elapsed, gps_index = 0, 0
        for index, image_name in enumerate(png_files):
            elapsed += 1 / fps
            writer.writerow([image_name, index, [gps[gps_index][0], gps[gps_index][1]]])
            if gps_index + 1 < len(gps) and elapsed >= gps[gps_index + 1][2]:
                gps_index += 1
