"""B200-native orthogroup-graph builder: a drop-in for OrthoFinder's WaterfallMethod hot path
(all-vs-all hit files -> OrthoFinder_graph.txt).  See DESIGN.md and include/ofg.h."""
__version__ = "0.1.0"
