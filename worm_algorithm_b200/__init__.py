"""worm_algorithm_b200 -- B200-native (sm_100a) worm-algorithm hot path behind the reference's
`WormSimulation` interface.  See DESIGN.md; the C ABI is include/worm_b200.h."""
__version__ = "0.1.0"
