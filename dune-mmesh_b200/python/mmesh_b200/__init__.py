"""mmesh_b200: B200-native moving-mesh hot path behind the Dune::MMesh adaptation API (Python test/bench binding)."""
