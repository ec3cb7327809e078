"""Host-side date arithmetic of the meteorology component: Julian day and the clock
time relative to true solar noon (TSN_offset) that ``tfb_met_energy`` takes as inputs.

Restates ``met_base.update_julian_day`` (components/met_base.py:1938-1993) and the
scalar helpers of ``components/solar_funcs.py`` it uses: ``Julian_Day`` :944-997,
``True_Solar_Noon`` :1347-1397, ``Equation_Of_Time`` :1214-1345, ``Earth_Perihelion``
:1122-1212, ``Vernal_Equinox`` :1092-1120.  Reference quirks kept on purpose: the hour
of day enters the Julian day as an INTEGER (minutes are dropped), and for a model year
outside 1981-2060 the perihelion date of the CURRENT calendar year is used.
"""
from datetime import datetime, timedelta

import numpy as np

DAYS_PER_YEAR = np.float64(365.2425)
ECCENTRICITY = np.float64(0.016713)
TILT_RAD = np.float64(23.4397) * (np.pi / np.float64(180))
SPIN_RATE = np.float64(2) * np.pi / np.float64(24)          # [rad / hour]

# (day of January, hour) of Earth's perihelion, 1981-2060 (solar_funcs.py:1140-1171)
_PERIHELION = dict(zip(range(1981, 2061), (
    (2, 2), (4, 11), (2, 15), (3, 22), (3, 20), (2, 5), (4, 23), (3, 0), (1, 22), (4, 17),
    (3, 3), (3, 15), (4, 3), (2, 6), (4, 11), (4, 7), (2, 0), (4, 21), (3, 13), (3, 5),
    (4, 9), (2, 14), (4, 5), (4, 18), (2, 1), (4, 15), (3, 20), (3, 0), (4, 15), (3, 0),
    (3, 19), (5, 0), (2, 5), (4, 12), (4, 7), (2, 23), (4, 14), (3, 6), (3, 5), (5, 8),
    (2, 14), (4, 7), (4, 16), (3, 1), (4, 13), (3, 17), (3, 3), (5, 12), (2, 18), (3, 10),
    (4, 21), (3, 5), (4, 12), (4, 5), (3, 1), (5, 14), (3, 4), (3, 5), (5, 7), (3, 12),
    (3, 22), (4, 9), (2, 22), (5, 13), (3, 15), (3, 1), (5, 12), (3, 18), (3, 10), (4, 20),
    (3, 6), (5, 9), (3, 22), (2, 18), (5, 12), (4, 4), (3, 3), (5, 4), (3, 11), (4, 23))))


def julian_day(month, day, hour=None, year=None):
    leap = (year is not None) and ((year % 4) == 0)
    month_days = np.array([0, 31, 29 if leap else 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
    JD = np.sum(month_days[:month]) + np.maximum(day - 1, 0)
    if hour is not None:
        JD = JD + (hour / np.float64(24))
    return JD


def earth_perihelion(year, current_year=None):
    if year < 1981 or year > 2060:
        year = datetime.now().year if current_year is None else int(current_year)
    d, h = _PERIHELION[int(year)]
    return julian_day(1, d, h)


def equation_of_time(JD, year, current_year=None):
    """[hours]"""
    twopi = np.float64(2) * np.pi
    Tp = earth_perihelion(year, current_year)
    M = (twopi / DAYS_PER_YEAR) * (JD - Tp)
    M = (M + twopi) % twopi
    VE = np.float64(79.3125) + DAYS_PER_YEAR * (year - np.float64(2000))
    PT = (np.float64(365) + Tp) - VE
    omega = twopi * (PT / DAYS_PER_YEAR)
    Lm = (M + omega)
    TE = (-2.0 * ECCENTRICITY * np.sin(M)) + (np.sin(2 * Lm) * (TILT_RAD / 2) ** 2.0)
    return TE / SPIN_RATE


def true_solar_noon(JD, lon_deg, GMT_offset, year, current_year=None):
    LC = ((GMT_offset * np.float64(15)) - lon_deg) / np.float64(15)
    return np.float64(12) + LC + equation_of_time(JD, year, current_year)


def day_and_tsn(start_datetime, time_min, lon_deg, GMT_offset, current_year=None):
    """(julian_day, TSN_offset [h]) for the model time `time_min` after
    `start_datetime` ('YYYY-MM-DD HH:MM:SS'); lon_deg may be a grid."""
    t = datetime.strptime(start_datetime, '%Y-%m-%d %H:%M:%S') + timedelta(minutes=float(time_min))
    JD = julian_day(t.month, t.day, hour=t.hour, year=t.year)
    clock_hour = (JD - np.int16(JD)) * np.float64(24)
    noon = true_solar_noon(JD, lon_deg, np.int16(GMT_offset), t.year, current_year)
    return JD, (clock_hour - noon)
