#!/usr/bin/env python
"""Mesh fixtures from the reference's own tutorial data (legacy DOLFIN XML, parseable without FEniCS):

    tutorials/01_thermal_block/data/thermal_block{,_physical_region,_facet_region}.xml   -> mesh_thermal_block.npz
    tutorials/02_elastic_block/data/elastic_block{,_physical_region,_facet_region}.xml   -> mesh_elastic_block.npz

Each .npz holds vertices (nv x 2), cells (nc x 3 vertex indices), cell_markers (nc) and the marked facets as rows
(cell, local facet, marker); DOLFIN's local facet i of a triangle is the edge OPPOSITE local vertex i.  The generator
reads /root/reference; the committed fixtures travel.  Run:  python tests/golden/make_mesh_fixtures.py"""
import os
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference/tutorials"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse(folder, name):
    base = os.path.join(REF, folder, "data", name)
    mesh = ET.parse(base + ".xml").getroot().find("mesh")
    verts = mesh.find("vertices")
    V = np.zeros((int(verts.get("size")), 2))
    for v in verts:
        V[int(v.get("index"))] = (float(v.get("x")), float(v.get("y")))
    cells = mesh.find("cells")
    Cc = np.zeros((int(cells.get("size")), 3), dtype=np.int64)
    for c in cells:
        Cc[int(c.get("index"))] = (int(c.get("v0")), int(c.get("v1")), int(c.get("v2")))
    markers = np.zeros(Cc.shape[0], dtype=np.int64)
    for val in ET.parse(base + "_physical_region.xml").getroot().iter("value"):
        markers[int(val.get("cell_index"))] = int(val.get("value"))
    facets = [(int(v.get("cell_index")), int(v.get("local_entity")), int(v.get("value")))
              for v in ET.parse(base + "_facet_region.xml").getroot().iter("value") if int(v.get("value")) != 0]
    return V, Cc, markers, np.array(facets, dtype=np.int64)


def main():
    for folder, name in (("01_thermal_block", "thermal_block"), ("02_elastic_block", "elastic_block")):
        V, Cc, markers, facets = parse(folder, name)
        np.savez_compressed(os.path.join(HERE, "mesh_%s.npz" % name), vertices=V, cells=Cc, cell_markers=markers,
                            marked_facets=facets)
        print(name, "vertices", V.shape[0], "cells", Cc.shape[0], "subdomains", sorted(set(markers.tolist())),
              "facet markers", sorted(set(facets[:, 2].tolist())))


if __name__ == "__main__":
    main()
