"""Host-framework plugins that call the hot path (bore/plugins/)."""
