"""CaDIS class tables the hot path needs: class ids per experiment and the mIoU category index lists.

Mirrors the *data* of the reference's utils/defaults.py:16-33 (categories) and :112-237 (class ids / names);
``CLASS_INFO[experiment][1]`` (id -> name, with 255 = "Ignore" for experiments 2 and 3) and
``CLASS_INFO[experiment][2]`` (category -> class ids) are indexed exactly like the reference's list.
Slot 0 of each entry (the raw-id remapping used by the dataset readers) is outside this path and left empty.
"""

_NAMES_1 = ["Pupil", "Surgical Tape", "Hand", "Eye Retractors", "Iris", "Skin", "Cornea", "Instrument"]
_NAMES_2 = ["Pupil", "Surgical Tape", "Hand", "Eye Retractors", "Iris", "Skin", "Cornea", "Cannula",
            "Cap. Cystotome", "Tissue Forceps", "Primary Knife", "Ph. Handpiece", "Lens Injector",
            "I/A Handpiece", "Secondary Knife", "Micromanipulator", "Cap. Forceps"]
_NAMES_3 = ["Pupil", "Surgical Tape", "Hand", "Eye Retractors", "Iris", "Skin", "Cornea", "Hydro. Cannula",
            "Visc. Cannula", "Cap. Cystotome", "Rycroft Cannula", "Bonn Forceps", "Primary Knife",
            "Ph. Handpiece", "Lens Injector", "I/A Handpiece", "Secondary Knife", "Micromanipulator",
            "I/A Handpiece Handle", "Cap. Forceps", "R. Cannula Handle", "Ph. Handpiece Handle",
            "Cap. Cystotome Handle", "Sec. Knife Handle", "Lens Injector Handle"]


def _classes(names, with_ignore):
    d = {i: n for i, n in enumerate(names)}
    if with_ignore:
        d[255] = "Ignore"
    return d


_ANATOMIES, _OTHERS = [0, 4, 5, 6], [1, 2, 3]
CATEGORIES = {
    0: {"anatomies": [], "instruments": [], "others": []},
    1: {"anatomies": _ANATOMIES, "instruments": [7], "others": _OTHERS, "rare": [2]},
    2: {"anatomies": _ANATOMIES, "instruments": list(range(7, 17)), "others": _OTHERS,
        "rare": [16, 10, 9, 12, 14]},
    3: {"anatomies": _ANATOMIES, "instruments": list(range(7, 25)), "others": _OTHERS,
        "rare": [24, 20, 21, 22, 18, 23, 19, 16, 12, 11, 14]},
}

CLASS_INFO = [
    [{}, {}, CATEGORIES[0]],
    [{}, _classes(_NAMES_1, False), CATEGORIES[1]],
    [{}, _classes(_NAMES_2, True), CATEGORIES[2]],
    [{}, _classes(_NAMES_3, True), CATEGORIES[3]],
]

NUM_CLASSES = {1: 8, 2: 17, 3: 25}          # network outputs; experiments 2/3 use label C as "ignore"


def mask_of(indices) -> int:
    m = 0
    for i in indices:
        if i != 255:
            m |= 1 << int(i)
    return m
