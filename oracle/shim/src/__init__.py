"""ORACLE (test infrastructure): import root `src` of the un-vendored USFlows package."""
