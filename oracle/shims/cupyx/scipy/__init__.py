'ORACLE SHIM — cupyx.scipy stand-in.'
