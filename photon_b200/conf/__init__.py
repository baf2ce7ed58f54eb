"""YAML tree of the configuration (``base.yaml`` + the ``llm_config`` / ``dataset`` / evaluation groups) and, for import-path parity
with the reference (``photon.conf.base_schema``), a module re-exporting the schema that lives in :mod:`photon_b200.config.schema`."""
