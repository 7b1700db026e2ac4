"""Lens prescriptions used as fixtures.

YAML accepted by the reference as ``System(**yaml.safe_load(text))``.  Only
tests/golden/make_golden.py (which needs the live reference) turns these into
packed surface tables; everything else reads the committed tables.
"""

# rayopt/test/test_raytrace.py:30-57 with Abbe-model glasses substituted for
# the catalog names (no sqlalchemy here): N-SK16 -> 1.62041/60.32,
# N-F2 -> 1.62005/36.43
COOKE = """
description: 'oslo cooke triplet example 50mm f/4 20deg'
wavelengths: [587.56e-9, 656.27e-9, 486.13e-9]
object: {angle_deg: 20, pupil: {radius: 6.25, aim: True}}
image: {type: finite, pupil: {radius: 0, update_radius: True}}
stop: 5
elements:
- {material: air}
- {roc: 21.25, distance: 5.0, material: 1.62041/60.32, radius: 6.5}
- {roc: -158.65, distance: 2.0, material: air, radius: 6.5}
- {roc: -20.25, distance: 6.0, material: 1.62005/36.43, radius: 5.0}
- {roc: 19.6, distance: 1.0, material: air, radius: 5.0}
- {material: air, radius: 4.75}
- {roc: 141.25, distance: 6.0, material: 1.62041/60.32, radius: 6.5}
- {roc: -17.285, distance: 2.0, material: air, radius: 6.5}
- {distance: 42.95, radius: 0.364}
"""

# C1: the README's 4-surface minimum (README.rst:124-126)
SINGLET = """
description: 'singlet'
object: {angle_deg: 5, pupil: {radius: 1}}
stop: 1
elements:
- {material: 1.0}
- {distance: 1, material: 1.5, roc: 5, radius: 1}
- {distance: .2, material: 1.0, roc: -5, radius: 1}
- {distance: 5, radius: 1}
"""

# C2 / C4: Double-Gauss, 12 traced surfaces
DOUBLE_GAUSS = """
description: 'double gauss 28 degree field'
wavelengths: [587.56e-9, 656.27e-9, 486.13e-9]
object: {angle_deg: 14, pupil: {radius: 16.67, aim: True}}
image: {type: finite, pupil: {radius: 0, update_radius: True}}
stop: 6
elements:
- {material: air}
- {distance: 10.0, roc: 54.153, material: 1.60738/56.65, radius: 29.225}
- {distance: 8.747, roc: 152.522, material: air, radius: 28.141}
- {distance: 0.5, roc: 35.951, material: 1.62041/60.32, radius: 24.296}
- {distance: 14.0, material: 1.60342/38.03, radius: 21.297}
- {distance: 3.777, roc: 22.270, material: air, radius: 14.919}
- {distance: 14.253, material: air, radius: 10.229}
- {distance: 12.428, roc: -25.685, material: 1.60342/38.03, radius: 13.188}
- {distance: 3.777, material: 1.62041/60.32, radius: 16.468}
- {distance: 10.834, roc: -36.980, material: air, radius: 18.930}
- {distance: 0.5, roc: 196.417, material: 1.62041/60.32, radius: 21.311}
- {distance: 6.858, roc: -67.148, material: air, radius: 21.646}
- {distance: 57.315, radius: 24.57}
"""

# C3: Cooke triplet with even aspheres on surfaces 1, 4, 7
COOKE_ASPH = """
description: 'cooke triplet with even aspheres'
wavelengths: [587.56e-9, 656.27e-9, 486.13e-9]
object: {angle_deg: 20, pupil: {radius: 6.25, aim: True}}
image: {type: finite, pupil: {radius: 0, update_radius: True}}
stop: 5
elements:
- {material: air}
- {roc: 21.25, distance: 5.0, material: 1.62041/60.32, radius: 6.5, aspherics: [0, 2.0e-6, -1.0e-8, 3.0e-11]}
- {roc: -158.65, distance: 2.0, material: air, radius: 6.5}
- {roc: -20.25, distance: 6.0, material: 1.62005/36.43, radius: 5.0}
- {roc: 19.6, distance: 1.0, material: air, radius: 5.0, conic: -0.3, aspherics: [0, -4.0e-6, 2.0e-8]}
- {material: air, radius: 4.75}
- {roc: 141.25, distance: 6.0, material: 1.62041/60.32, radius: 6.5}
- {roc: -17.285, distance: 2.0, material: air, radius: 6.5, aspherics: [0, 1.5e-6, 0, 0]}
- {distance: 42.95, radius: 20}
"""

# C5 / headline: zoom-like, 20 traced surfaces
ZOOM = """
description: 'zoom-like 20 surface'
wavelengths: [587.56e-9, 656.27e-9, 486.13e-9, 546.07e-9, 435.84e-9]
object: {angle_deg: 10, pupil: {radius: 12.0, aim: True}}
image: {type: finite, pupil: {radius: 0, update_radius: True}}
stop: 6
elements:
- {material: air}
- {distance: 10.0, roc: 54.153, material: 1.60738/56.65, radius: 29.225}
- {distance: 8.747, roc: 152.522, material: air, radius: 28.141}
- {distance: 0.5, roc: 35.951, material: 1.62041/60.32, radius: 24.296}
- {distance: 14.0, material: 1.60342/38.03, radius: 21.297}
- {distance: 3.777, roc: 22.270, material: air, radius: 14.919}
- {distance: 14.253, material: air, radius: 10.229}
- {distance: 12.428, roc: -25.685, material: 1.60342/38.03, radius: 13.188}
- {distance: 3.777, material: 1.62041/60.32, radius: 16.468}
- {distance: 10.834, roc: -36.980, material: air, radius: 18.930}
- {distance: 0.5, roc: 196.417, material: 1.62041/60.32, radius: 21.311}
- {distance: 6.858, roc: -67.148, material: air, radius: 21.646}
- {distance: 4.0, roc: 120, material: 1.51680/64.17, radius: 22}
- {distance: 4.0, roc: -300, material: air, radius: 22}
- {distance: 6.0, roc: -90, material: 1.62005/36.43, radius: 21}
- {distance: 2.5, roc: 250, material: air, radius: 21}
- {distance: 8.0, roc: 80, material: 1.51680/64.17, radius: 22}
- {distance: 5.0, roc: -400, material: air, radius: 22}
- {distance: 3.0, roc: 60, material: 1.62041/60.32, radius: 21}
- {distance: 4.5, roc: 45, material: air, radius: 19}
- {distance: 30, radius: 25}
"""

# rotated frames + mirror + negative distance: rayopt/test/test_seidel.py:27-39
# with conic 0 (the conic -1 original gives 0/0 on axis)
MIRROR = """
description: 'spherical mirror, folded (negative distance flips the frame)'
object: {type: infinite, angle_deg: 1, pupil: {radius: 1, distance: 1}}
stop: 1
elements:
- {material: vacuum}
- {material: mirror, distance: 1, roc: -200, radius: 5}
- {material: vacuum, distance: -100, radius: 5}
"""

SYSTEMS = {
    "cooke": COOKE,
    "singlet": SINGLET,
    "double_gauss": DOUBLE_GAUSS,
    "cooke_asph": COOKE_ASPH,
    "zoom": ZOOM,
    "mirror": MIRROR,
}
