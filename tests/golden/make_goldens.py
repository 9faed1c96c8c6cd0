"""Writes tests/golden/reference_goldens.json.

The reference cannot be built or imported in this container, so these are the values its OWN
tests assert, transcribed from the test sources (file:line given per entry), plus the hand-reconstructible
fixtures those tests run on (the .msh files are git-LFS pointers; the .geo sources are in the tree).
Run:  python tests/golden/make_goldens.py
"""
import json
import os

G = {
    "simple2d": {
        "source": "dune/mmesh/test/grids/simple2d.geo:3-8,27-31; asserted in dune/mmesh/test/test-mmesh-2d.cc:139-223",
        "vertices": [[0, 0], [1, 0], [1, 1], [0, 1], [1, 0.25], [0.5, 1]],
        "vertex_ids": [0, 1, 2, 3, 4, 5],
        # element order after the x-centroid sort of explicitgridfactory.hh:391-403
        "triangles": [[5, 3, 0], [0, 4, 5], [0, 1, 4], [4, 2, 5]],
        "sizes": {"elements": 4, "edges": 9, "vertices": 6},
        "elements": {
            "1": {"center": [0.5, 0.41666666666666663], "volume": 0.4375, "sub_indices": [0, 4, 5],
                  "edge_centers": [[0.5, 0.125], [0.25, 0.5], [0.75, 0.625]]},
            "2": {"center": [0.66666666666666663, 0.083333333333333329], "volume": 0.125, "sub_indices": [0, 1, 4],
                  "edge_centers": [[0.5, 0], [0.5, 0.125], [1, 0.125]]},
        },
    },
    "twocell3d": {
        "source": "dune/mmesh/test/test-mmesh-3d.cc:133-193",
        "vertices": [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1]],
        "tets": [[0, 1, 3, 4], [1, 2, 3, 4]],
        "elements": {
            "0": {"center": [0.375, 0.375, 0.25], "volume": 1.0 / 6.0, "sub_indices": [0, 1, 3, 4]},
            "1": {"center": [0.625, 0.625, 0.25], "volume": 0.16666666666666666},
        },
    },
    "tolerances": {
        "distance_2d_rel": {"value": 1e-8, "source": "dune/mmesh/test/test-distance.cc:18-37,57"},
        "distance_3d_rel": {"value": 0.05, "source": "dune/mmesh/test/test-distance.cc:63"},
        "distance_integral": {"value": 0.25, "tol": 1e-8, "source": "python/dune/mmesh/test/distance.py:24-28"},
        "clip_area_abs": {"value": 1e-10, "source": "dune/mmesh/test/test-intersectionvolume.cc:52-61"},
        "conservation_abs": {"value": 1e-8, "source": "dune/mmesh/test/examples/movement.cc:89-94"},
        "projection_l2": {"value": 1e-12, "source": "dune/mmesh/test/test-femadaptation.cc:286-287"},
    },
    "indicator_params_femadaptation": {
        "source": "dune/mmesh/test/test-femadaptation.cc:279-281,219-220",
        "maxH": 0.75, "minH": 0.05, "factor": 1.0, "scale_maxH": 0.8, "scale_minH": 1.2,
    },
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    print("wrote", out)
