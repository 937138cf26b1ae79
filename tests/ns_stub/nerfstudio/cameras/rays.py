"""The stub does not replace the ray containers: raising here makes umhs_b200.rays keep its own stand-ins."""
raise ImportError("tests/ns_stub has no nerfstudio.cameras.rays")
