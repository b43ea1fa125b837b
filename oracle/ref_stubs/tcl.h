/* opaque stand-ins so the reference's TclCommands.h parses (oracle/_ref build only) */
#ifndef STUB_TCL_H
#define STUB_TCL_H
typedef void *ClientData;
struct Tcl_Interp;
struct Tcl_Obj;
#endif
