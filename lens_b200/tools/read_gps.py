"""NMEA track reader (mirror of lens/tools/read_gps.py:5-31): rows of (latitude, longitude, seconds since
the first fix), keeping only fixes that moved by more than 1e-4 degrees.  Needs pynmea2, like the reference."""
import numpy as np


def get_gps(nmea_file_path):
    import pynmea2
    rows, first, prev = [], None, (0, 0)
    with open(nmea_file_path, encoding="utf-8") as fh:
        for line in fh:
            try:
                msg = pynmea2.parse(line)
            except pynmea2.ParseError:
                continue
            if first is None:
                first = msg.timestamp
            if msg.sentence_type in ("GSV", "VTG", "GSA"):
                continue
            lat, lon = msg.latitude, msg.longitude
            moved = np.linalg.norm(np.array([lat, lon]) - np.array(prev))
            if lat != 0 and lon != 0 and lat != prev[0] and lon != prev[1] and moved > 0.0001:
                ts = msg.timestamp
                rows.append((lat, lon, (ts.hour - first.hour) * 3600 + (ts.minute - first.minute) * 60
                             + (ts.second - first.second)))
                prev = (lat, lon)
    return np.array(rows, dtype=np.float64).reshape(-1, 3)
